#!/bin/bash
# Round 2, call Z (8 GPUs): bench.py as the driver launches it at N=8 (all rows), then the 2-rank stitch tests.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L | wc -l > gpurun_out/r2z_gpus.txt
echo "== bench --gpus 8 (all rows)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/r2z_bench8.err | tee gpurun_out/r2z_bench_n8.json | cut -c1-300
tail -6 gpurun_out/r2z_bench8.err | cut -c1-300
echo "== bench --gpus 4 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29822 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/r2z_bench4.err | tee gpurun_out/r2z_bench_n4.json | cut -c1-300
tail -4 gpurun_out/r2z_bench4.err | cut -c1-300
