mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*umma -s 6 -c 2 -f -o gpurun_out/prof_attn32 python bench.py --workload attn --steps 1 --warmup 2 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/ncu32.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -2 gpurun_out/ncu32.log | cut -c1-200; wc -l gpurun_out/launches.csv
