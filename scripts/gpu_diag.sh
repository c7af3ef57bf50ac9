# Diagnostic lease: tcgen05 building blocks, mixed-format probe (own process), parity suite with full log, headline bench,
# backward timeline (HSTU_TRACE rebuild), one Triton comparator shape with the full error text.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 300 python bench.py --workload attn --batch 512 --attn-heads 4 --attn-dim 64 --lmax 2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn64.log 2>&1
timeout 300 python bench.py --workload attn --batch 512 --attn-heads 4 --attn-dim 128 --lmax 2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn128.log 2>&1

HSTU_EXP="HSTU_TRACE" timeout 400 python -m generative_recommenders_b200.build --force > gpurun_out/trace_build.log 2>&1
bash scripts/gpu_trace.sh > gpurun_out/trace_run.log 2>&1
python scripts/trace_report.py gpurun_out/bwd_trace.txt 10 4 > gpurun_out/bwd_timeline.txt 2>&1
echo "== pytest"; tail -12 gpurun_out/pytest_gpu_full.log | cut -c1-330
echo "== bench"; tail -1 gpurun_out/bench_default.log | cut -c1-2500
echo "== attn64"; tail -1 gpurun_out/bench_attn64.log | cut -c1-1500
echo "== attn128"; tail -1 gpurun_out/bench_attn128.log | cut -c1-1500

echo "== timeline"; head -70 gpurun_out/bwd_timeline.txt
