mkdir -p gpurun_out
HSTU_EXP="HSTU_TRACE HSTU_EXP_NO_ELEM" python -m generative_recommenders_b200.build --force > gpurun_out/exp_build.log 2>&1
timeout 120 python scripts/dbg_bwd.py 2 8192 8 > gpurun_out/dbg_trace.log 2>&1
tail -2 gpurun_out/dbg_trace.log; wc -l gpurun_out/bwd_trace.txt
