mkdir -p gpurun_out
timeout 120 python scripts/dbg_bwd.py 2 8192 8 > gpurun_out/dbg_trace.log 2>&1
tail -3 gpurun_out/dbg_trace.log; wc -l gpurun_out/bwd_trace.txt
