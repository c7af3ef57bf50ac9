# final tree: whole GPU suite, smoke, headline bench, our attention sweep (the Triton lines of the earlier run are reused for the table)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; tail -3 gpurun_out/pytest_gpu_full.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-400
cp profiles/r02_attn_sweep_triton.jsonl gpurun_out/sweep_triton.jsonl
bash scripts/gpu_triton_sweep.sh ours
