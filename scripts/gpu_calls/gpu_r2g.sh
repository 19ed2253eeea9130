#!/bin/bash
# Round 2, call G (2 GPUs): the device-side stitch test and bench.py under torchrun.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/r2g_gpus.txt
echo "== stitch test"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "device_stitch or multi_gpu" 2>&1 | tail -25 | tee gpurun_out/r2g_pytest.log
echo "== bench --gpus 2 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2g_bench2.err | tee gpurun_out/r2g_bench_n2_headline.json | cut -c1-900
tail -15 gpurun_out/r2g_bench2.err | cut -c1-300
echo "== bench --gpus 2 (all configurations)"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2g_bench2b.err | tee gpurun_out/r2g_bench_n2.json | cut -c1-600
tail -15 gpurun_out/r2g_bench2b.err | cut -c1-300
