# persistent forward at d = 128: parity (attention + full-size files), then A/B against the one-CTA-per-item kernel via HSTU_FWD_PERSIST
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity_fullsize.py tests/test_gpu_attention.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
: > gpurun_out/ab_persist.txt
for shape in "512 4 128 512" "512 4 128 2048" "512 4 64 512"; do
  set -- $shape
  for pv in 0 1; do
    r=$(HSTU_FWD_PERSIST=$pv timeout 300 python bench.py --workload attn --batch $1 --attn-heads $2 --attn-dim $3 --lmax $4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('fwd %.4f bwd %.4f'%(r['fwd']['ms_per_launch'], r['ms_per_launch']))")
    echo "[persist=$pv] B=$1 H=$2 d=$3 L=$4: $r" | tee -a gpurun_out/ab_persist.txt
  done
done
