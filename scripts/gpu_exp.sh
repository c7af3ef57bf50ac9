mkdir -p gpurun_out
for exp in "HSTU_EXP_NO_ELEM"; do
  HSTU_EXP="$exp" python -m generative_recommenders_b200.build --force > gpurun_out/exp_build.log 2>&1
  echo "== exp [$exp]"
  timeout 300 python bench.py --workload attn --steps 5 --warmup 3 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline 2>&1 | tail -1 | grep -o "kernel_ms_per_call.*\|rror.*" | cut -c1-200
  timeout 300 python bench.py --workload attn --steps 5 --warmup 3 --batch 32 --lmax 4096 --attn-dim 128 --no-cpu-baseline 2>&1 | tail -1 | grep -o "kernel_ms_per_call.*\|rror.*" | cut -c1-200
done
