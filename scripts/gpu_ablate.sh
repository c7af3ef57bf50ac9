# timing-only ablations (wrong numerics) of the attention kernels on the headline shape: build the variants first, e.g.
#   python scripts/build_variant.py noelem "HSTU_EXP_BWD_NO_ELEM HSTU_EXP_NO_ELEM" attn_umma_bwd.cu,attn_umma_fwd.cu
#   python scripts/build_variant.py nomufu HSTU_EXP_NO_MUFU attn_umma_bwd.cu,attn_umma_fwd.cu
#   python scripts/build_variant.py nodrain HSTU_EXP_BWD_NO_DRAIN attn_umma_bwd.cu
#   gpurun -- bash scripts/gpu_ablate.sh noelem nomufu nodrain
mkdir -p gpurun_out
: > gpurun_out/ablate.txt
for v in default "$@"; do
  if [ "$v" = default ]; then unset HSTU_B200_LIB; else export HSTU_B200_LIB=$PWD/generative_recommenders_b200/lib/variants/libhstu_b200_$v.so; fi
  r=$(timeout 300 python bench.py --workload attn --batch 16 --attn-heads 8 --attn-dim 32 --lmax 8192 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.3f bwd %.3f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
  echo "[$v] d=32: $r" | tee -a gpurun_out/ablate.txt
done
