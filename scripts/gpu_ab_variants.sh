# A/B + parity of forward variants: headline shape and the reference's short / medium microbenchmark shapes
mkdir -p gpurun_out
: > gpurun_out/ab.txt
for v in default "$@"; do
  if [ "$v" = default ]; then unset HSTU_B200_LIB; else export HSTU_B200_LIB=$PWD/generative_recommenders_b200/lib/variants/libhstu_b200_$v.so; fi
  if [ "$v" != default ]; then
    timeout 600 python -m pytest tests/test_gpu_parity_fullsize.py tests/test_gpu_attention.py tests/test_gpu_block.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
  fi
  for shape in "16 8 32 8192" "512 4 64 512" "512 4 64 2048" "512 4 32 512" "128 8 32 256"; do
    set -- $shape
    r=$(timeout 300 python bench.py --workload attn --batch $1 --attn-heads $2 --attn-dim $3 --lmax $4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.4f bwd %.4f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
    echo "[$v] B=$1 H=$2 d=$3 L=$4: $r" | tee -a gpurun_out/ab.txt
  done
done
