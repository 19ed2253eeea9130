#!/bin/bash
# Round 2, call AI (8 GPUs): the headline at N=8 and N=4 with the fence-free stitch, full rows at N=8.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== bench --gpus 8 (all rows)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/r2ai_bench8.err | tee gpurun_out/r2ai_bench_n8.json | cut -c1-200
tail -3 gpurun_out/r2ai_bench8.err | cut -c1-300
echo "== bench --gpus 4 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29832 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/r2ai_bench4.err | tee gpurun_out/r2ai_bench_n4.json | cut -c1-200
echo "== bench --gpus 2 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2ai_bench2.err | tee gpurun_out/r2ai_bench_n2.json | cut -c1-200
echo "== bench N=1 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2ai_bench1.err | tee gpurun_out/r2ai_bench_n1.json | cut -c1-200
