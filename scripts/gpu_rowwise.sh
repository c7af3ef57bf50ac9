# Row-wise kernels: microbenchmark of the product library and of the variants given as arguments (names under lib/variants),
# one ncu full-set capture of each row-wise kernel, the relative-bias attention tests and the ML-20M bench line.
mkdir -p gpurun_out
echo "[default] $(timeout 300 python scripts/rowwise_bench.py 2>&1 | tail -1)" | tee gpurun_out/rowwise.txt
for v in "$@"; do
  echo "[$v] $(HSTU_B200_LIB=$PWD/generative_recommenders_b200/lib/variants/libhstu_b200_$v.so timeout 300 python scripts/rowwise_bench.py 2>&1 | tail -1)" | tee -a gpurun_out/rowwise.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd|ln_bwd|nmd_fwd|nmd_bwd|silu_kernel|colsum" -c 60 -f -o gpurun_out/prof_rowwise_r02 python scripts/rowwise_bench.py --iters 2 > gpurun_out/ncu_rowwise.log 2>&1
ls -la gpurun_out/prof_rowwise_r02.ncu-rep
timeout 600 python -m pytest tests/test_gpu_research_block.py tests/test_gpu_attention.py -m gpu -q -x -k "bias or research or cache" -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --workload ml20m 2>&1 | tail -1 | cut -c1-1800
