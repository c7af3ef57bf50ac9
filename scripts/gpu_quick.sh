# quick lease: selftest report (conversion / MUFU rates) + attention microbenches + headline bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_umma.py -q --tb=short -p no:cacheprovider -k selftest > gpurun_out/selftest_pytest.log 2>&1
timeout 300 python -m pytest tests/test_gpu_attention.py tests/test_gpu_parity_fullsize.py -q --tb=line -p no:cacheprovider -x > gpurun_out/pytest_attn.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_default.log 2>&1
timeout 300 python bench.py --workload attn --batch 512 --attn-heads 4 --attn-dim 64 --lmax 2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn64.log 2>&1
timeout 300 python bench.py --workload attn --batch 512 --attn-heads 4 --attn-dim 128 --lmax 2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_attn128.log 2>&1
grep -E "^mufu|cvt|fma" gpurun_out/umma_selftest.txt
tail -3 gpurun_out/pytest_attn.log | cut -c1-300
python - <<'PY'
import json
for f in ['bench_default','bench_attn64','bench_attn128']:
    try:
        l=open(f'gpurun_out/{f}.log').read().strip().split('\n')[-1]
        d=json.loads(l); r=d['roofline']
        print(f, 'ms/step %.2f'%d['ms_per_step'], 'fwd %.3f ms (%.3f) bwd %.3f ms (%.3f)'%(r['fwd']['ms_per_launch'], r['fwd']['frac'], r['ms_per_launch'], r['frac']), d.get('kernel_ms_per_call'))
    except Exception as e: print(f, 'ERR', e, l[-300:])
PY
