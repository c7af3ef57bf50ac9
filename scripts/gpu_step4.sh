# fast row-wise kernels: parity, microbench, whole suite, headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_norm_fast.py tests/test_gpu_block.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
echo "[fast] $(timeout 300 python scripts/rowwise_bench.py 2>&1 | tail -1)" | tee gpurun_out/rowwise_fast.txt
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; tail -3 gpurun_out/pytest_gpu_full.log
timeout 500 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-3000
