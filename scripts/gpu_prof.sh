mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_.*umma -s 6 -c 2 -o gpurun_out/prof_attn32 python bench.py --workload attn --steps 1 --warmup 2 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/ncu32.log 2>&1
tail -2 gpurun_out/ncu32.log | cut -c1-300
