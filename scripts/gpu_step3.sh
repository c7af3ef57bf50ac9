# ncu of the row-wise kernels (one launch each, small report), A/B of backward variants given as arguments, research bench lines
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:"ln_fwd|ln_bwd|nmd_fwd|nmd_bwd|silu_kernel|colsum" -c 12 -f -o gpurun_out/prof_rowwise_r02 python scripts/rowwise_bench.py --profile --sets 1 > gpurun_out/ncu_rowwise.log 2>&1
ls -la gpurun_out/prof_rowwise_r02.ncu-rep
: > gpurun_out/ab.txt
for v in default "$@"; do
  if [ "$v" = default ]; then unset HSTU_B200_LIB; else export HSTU_B200_LIB=$PWD/generative_recommenders_b200/lib/variants/libhstu_b200_$v.so; fi
  r=$(timeout 300 python bench.py --workload attn --batch 16 --attn-heads 8 --attn-dim 32 --lmax 8192 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.3f bwd %.3f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
  echo "[$v] d=32: $r" | tee -a gpurun_out/ab.txt
  if [ "$v" != default ]; then timeout 300 python -m pytest tests/test_gpu_parity_fullsize.py -m gpu -q -x -k "32" -p no:cacheprovider 2>&1 | tail -1; fi
done
unset HSTU_B200_LIB
for w in ml20m amzn_books; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/bench_$w.json; cut -c1-400 gpurun_out/bench_$w.json; done
