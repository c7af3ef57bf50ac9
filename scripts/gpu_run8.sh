mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_attention.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -12 > gpurun_out/pytest8.log
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/bench_attn32.log 2>&1
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 64 --lmax 2048 --attn-dim 64 --no-cpu-baseline > gpurun_out/bench_attn64.log 2>&1
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 32 --lmax 4096 --attn-dim 128 --no-cpu-baseline > gpurun_out/bench_attn128.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_large.log 2>&1
for f in pytest8 bench_attn32 bench_attn64 bench_attn128 bench_large; do echo "== $f"; tail -n 3 gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    try:
        d=json.loads(l); r=d['roofline']; print(d['config']['workload'][:60], 'ms/step',round(d['ms_per_step'],3),'fwd TF',round(r['fwd']['achieved'],1),'ms',round(r['fwd']['ms_per_launch'],3),'bwd TF',round(r['achieved'],1),'ms',round(r['ms_per_launch'],3), d.get('kernel_ms_per_call'))
    except Exception as e: print(l[:300])
"; done
