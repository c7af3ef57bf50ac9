mkdir -p gpurun_out
timeout 120 python scripts/dbg_bwd.py 2 8192 8 > gpurun_out/dbg_trace.log 2>&1
tail -3 gpurun_out/dbg_trace.log
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['kernel_ms_per_call'])"
