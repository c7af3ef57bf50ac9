# backward timeline of the current kernel (HSTU_TRACE rebuild on the box), d = 32
mkdir -p gpurun_out
HSTU_EXP="HSTU_TRACE" timeout 400 python -m generative_recommenders_b200.build --force > gpurun_out/trace_build.log 2>&1
bash scripts/gpu_trace.sh > gpurun_out/trace_run.log 2>&1
python scripts/trace_report.py gpurun_out/bwd_trace.txt 20 8 > gpurun_out/bwd_timeline.txt 2>&1
head -120 gpurun_out/bwd_timeline.txt
