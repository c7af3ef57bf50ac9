# One GPU lease: parity suite + smoke + headline bench.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1
for f in pytest_gpu smoke; do echo "== $f"; tail -n 25 gpurun_out/$f.log | cut -c1-400; done
echo "== bench"; tail -1 gpurun_out/bench_default.log | cut -c1-3500
