mkdir -p gpurun_out
for exp in "HSTU_EXP_NO_MUFU" "HSTU_EXP_NO_STS" "HSTU_EXP_NO_MUFU HSTU_EXP_NO_STS"; do
  HSTU_EXP="$exp" python -m generative_recommenders_b200.build --force > /dev/null 2>&1
  echo "== exp [$exp]"
  timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_call'])"
done
