# Full check of one lease: whole GPU parity suite, smoke, headline bench, attention sweep of ours; then the Triton comparator.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1
echo "== pytest"; tail -8 gpurun_out/pytest_gpu_full.log | cut -c1-300
echo "== smoke"; tail -2 gpurun_out/smoke.log | cut -c1-300
echo "== bench"; tail -1 gpurun_out/bench_default.log | cut -c1-2000
bash scripts/gpu_triton_sweep.sh both
