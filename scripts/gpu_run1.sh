mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 300 python -c "
import ctypes as C, sys
sys.path.insert(0,'.')
from generative_recommenders_b200 import _lib
buf=C.create_string_buffer(1<<16)
r=_lib.lib().hstu_umma_selftest(buf,len(buf))
open('gpurun_out/selftest.txt','w').write(buf.value.decode()+'\nrc=%d\n'%r)
print(buf.value.decode(), r)
" > gpurun_out/selftest.log 2>&1
timeout 900 python -m pytest tests/test_gpu_block.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/pytest_block.log
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -120 > gpurun_out/pytest_attn.log
timeout 600 python -m pytest tests/test_gpu_umma.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/pytest_umma.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --batch 2 --layers 2 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1
timeout 600 python bench.py --workload attn --steps 3 --warmup 2 --batch 32 --lmax 2048 --attn-dim 64 > gpurun_out/bench_attn64.log 2>&1
tail -5 gpurun_out/selftest.log gpurun_out/pytest_block.log gpurun_out/pytest_attn.log gpurun_out/pytest_umma.log gpurun_out/smoke.log gpurun_out/bench_small.log gpurun_out/bench_attn64.log
