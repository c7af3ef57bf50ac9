# relative bias on the tcgen05 kernels: parity, the research-path bench lines, and the headline attention shape (must be unchanged)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_research_block.py tests/test_gpu_parity_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
r=$(timeout 300 python bench.py --workload attn --batch 16 --attn-heads 8 --attn-dim 32 --lmax 8192 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.3f bwd %.3f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
echo "[headline attention] $r"
for w in ml20m amzn_books; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_$w.json; cut -c1-100 gpurun_out/bench_$w.json; python -c "import json; d=json.loads(open('gpurun_out/bench_$w.json').read()); print(round(d['value'],1), round(d['ms_per_step'],2), d['kernel_ms_per_call'])"; done
