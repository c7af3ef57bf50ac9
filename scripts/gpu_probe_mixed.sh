# Probe (own process: the error poisons the CUDA context): does kind::f16 accept an fp16 A operand with a bf16 B operand?
mkdir -p gpurun_out
HSTU_SELFTEST_MIXED=1 timeout 120 python - > gpurun_out/mixed_probe.log 2>&1 <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
from generative_recommenders_b200 import _lib
buf = C.create_string_buffer(1 << 16)
rc = _lib.selftest_lib().hstu_umma_selftest(buf, len(buf))
print(buf.value.decode()[-1500:])
print('rc', rc)
PY
tail -8 gpurun_out/mixed_probe.log
