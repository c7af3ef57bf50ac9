# The GPU comparator next to our kernels, same seeded inputs, same CUDA-event method (bench.py --workload attn):
#   headline attention shape (B=16, H=8, d=32, Lmax=8192) and the BASELINE microbench grid B=512, H=4, d in {64,128}, Lmax in {512,2048,8192}.
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_triton_sweep.sh [ours|triton|both]'
mkdir -p gpurun_out
WHAT=${1:-both}
SHAPES="16 8 32 8192|512 4 64 2048|512 4 64 8192|512 4 64 512|512 4 128 2048|512 4 128 8192|512 4 128 512"
if [ "$WHAT" != "triton" ]; then
  : > gpurun_out/sweep_ours.jsonl
  IFS='|'; for cfg in $SHAPES 512\ 4\ 256\ 512 512\ 4\ 256\ 2048; do IFS=' '; set -- $cfg
    timeout 200 python bench.py --workload attn --steps 3 --warmup 3 --batch $1 --attn-heads $2 --attn-dim $3 --lmax $4 --no-cpu-baseline 2>gpurun_out/sweep_ours.err | tail -1 >> gpurun_out/sweep_ours.jsonl
  IFS='|'; done; IFS=' '
fi
if [ "$WHAT" != "ours" ]; then
  : > gpurun_out/sweep_triton.jsonl
  IFS='|'; for cfg in $SHAPES; do IFS=' '; set -- $cfg
    timeout 900 python bench.py --impl triton --workload attn --steps 3 --warmup 2 --batch $1 --attn-heads $2 --attn-dim $3 --lmax $4 2>gpurun_out/sweep_triton.err | tail -1 >> gpurun_out/sweep_triton.jsonl
  IFS='|'; done; IFS=' '
fi
python scripts/sweep_table.py gpurun_out/sweep_ours.jsonl gpurun_out/sweep_triton.jsonl | tee gpurun_out/sweep_table.txt
