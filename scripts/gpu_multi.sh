# 2 GPUs of one box: weak scaling (per-GPU work fixed), strong scaling (global batch fixed, sharded by cost), the reference arm under torchrun
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
tail -1 gpurun_out/bench_2gpu.log | cut -c1-1500
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > gpurun_out/bench_2gpu_strong.log 2>&1
tail -1 gpurun_out/bench_2gpu_strong.log | cut -c1-1500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_ref.log 2>&1
tail -1 gpurun_out/bench_2gpu_ref.log | cut -c1-600
