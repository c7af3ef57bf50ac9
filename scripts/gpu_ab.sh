# A/B of compile-time variants of the backward kernel on the headline attention shape: HSTU_EXP defines, rebuilt on the box
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for exp in "$@"; do
  HSTU_EXP="$exp" timeout 400 python -m generative_recommenders_b200.build --force > gpurun_out/ab_build.log 2>&1
  r=$(timeout 300 python bench.py --workload attn --batch ${AB_B:-16} --attn-heads ${AB_H:-8} --attn-dim ${AB_D:-32} --lmax ${AB_L:-8192} --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.3f bwd %.3f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
  echo "[$exp] d=32: $r" | tee -a gpurun_out/ab.txt
done
