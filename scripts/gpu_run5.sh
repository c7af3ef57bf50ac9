mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_attention.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -8 > gpurun_out/pytest5.log
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/bench_attn32.log 2>&1
timeout 600 python bench.py --workload attn --steps 5 --warmup 3 --batch 64 --lmax 2048 --attn-dim 64 --no-cpu-baseline > gpurun_out/bench_attn64.log 2>&1
( time timeout 1200 python bench.py ) > gpurun_out/bench_default.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
for f in pytest5 bench_attn32 bench_attn64; do echo "== $f"; tail -n 2 gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    try:
        d=json.loads(l); r=d['roofline']; print(d['config']['workload'][:60], 'ms/step',round(d['ms_per_step'],3),'fwd TF',round(r['fwd']['achieved'],1),'ms',round(r['fwd']['ms_per_launch'],3),'bwd TF',round(r['achieved'],1),'ms',round(r['ms_per_launch'],3))
    except Exception as e: print(l[:300])
"; done
tail -n 6 gpurun_out/bench_default.log | cut -c1-2500
tail -n 5 gpurun_out/bench_ref.log | cut -c1-1200
