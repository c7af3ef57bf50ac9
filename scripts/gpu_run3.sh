mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_attention.py tests/test_gpu_block.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest3.log
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 64 --lmax 2048 --attn-dim 64 --no-cpu-baseline > gpurun_out/bench_attn64.log 2>&1
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/bench_attn32.log 2>&1
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 32 --lmax 4096 --attn-dim 128 --no-cpu-baseline > gpurun_out/bench_attn128.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 --batch 16 --no-cpu-baseline > gpurun_out/bench_large.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_.*umma -s 6 -c 2 -o gpurun_out/prof_attn32 python bench.py --workload attn --steps 1 --warmup 2 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/ncu32.log 2>&1
for f in smoke pytest3 bench_attn64 bench_attn32 bench_attn128 bench_large; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-1500; done
