# TMEM read-rate microbenchmark + fp16-accumulator probe (own processes), and the ablation "skip every other TMEM load" on the attention kernels
mkdir -p gpurun_out
HSTU_SELFTEST_F16ACC=1 timeout 120 python - > gpurun_out/f16acc_probe.log 2>&1 <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from generative_recommenders_b200 import _lib
buf = C.create_string_buffer(1 << 16)
rc = _lib.selftest_lib().hstu_umma_selftest(buf, len(buf))
print(buf.value.decode()[-2500:])
print('rc', rc)
PY
tail -8 gpurun_out/f16acc_probe.log
timeout 300 python - > gpurun_out/selftest_full.log 2>&1 <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from generative_recommenders_b200 import _lib
buf = C.create_string_buffer(1 << 16)
rc = _lib.selftest_lib().hstu_umma_selftest(buf, len(buf))
print(buf.value.decode())
print('rc', rc)
PY
grep -E "tmem-rate|failed checks|rc " gpurun_out/selftest_full.log
: > gpurun_out/ab.txt
for v in default "$@"; do
  if [ "$v" = default ]; then unset HSTU_B200_LIB; else export HSTU_B200_LIB=$PWD/generative_recommenders_b200/lib/variants/libhstu_b200_$v.so; fi
  r=$(timeout 300 python bench.py --workload attn --batch 16 --attn-heads 8 --attn-dim 32 --lmax 8192 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.3f bwd %.3f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
  echo "[$v] d=32: $r" | tee -a gpurun_out/ab.txt
done
echo "[rowwise] $(timeout 300 python scripts/rowwise_bench.py --only silu_fwd,silu_bwd 2>&1 | tail -1)"
