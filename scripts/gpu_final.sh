mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1
for f in pytest_final smoke; do echo "== $f"; tail -n 3 gpurun_out/$f.log | cut -c1-300; done
echo "== bench"; tail -1 gpurun_out/bench_default.log | cut -c1-3000
echo "== ref"; tail -1 gpurun_out/bench_ref.log | cut -c1-900
