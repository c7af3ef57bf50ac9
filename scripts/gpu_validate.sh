# One lease: whole GPU suite + smoke + headline bench on the current tree, the two research-path workloads, the reference arm,
# then the ncu launch list of the bench step and the full-set capture of the attention kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1
timeout 500 python bench.py > gpurun_out/bench_default.log 2>&1
echo "== pytest"; tail -12 gpurun_out/pytest_gpu_full.log | cut -c1-300
echo "== smoke"; tail -2 gpurun_out/smoke.log | cut -c1-300
echo "== bench"; tail -1 gpurun_out/bench_default.log | cut -c1-3000
for w in ml20m amzn_books; do
  timeout 300 python bench.py --workload $w > gpurun_out/bench_$w.log 2>&1
  echo "== $w"; tail -3 gpurun_out/bench_$w.log | cut -c1-1800
done
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.log 2>&1
echo "== reference arm"; tail -1 gpurun_out/bench_reference.log | cut -c1-1200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lb.log 2>&1
echo "== launch list rows"; wc -l gpurun_out/launches_bench.csv
bash scripts/gpu_ncu.sh
