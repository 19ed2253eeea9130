#!/bin/bash
# N-GPU bench exactly as the driver launches it (torchrun, one rank per GPU).
set -u
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== bench --gpus $N"
timeout ${RJ_MULTI_TIMEOUT:-150} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_ours_n$N.json | cut -c1-900
echo "rc=$?"; grep -vE "^\s*$|OMP_NUM_THREADS|\*\*\*\*" gpurun_out/bench_n$N.err | tail -12 | cut -c1-300
if [ "${RJ_REF:-0}" = "1" ]; then
echo "== reference --gpus $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>> gpurun_out/bench_n$N.err | tee gpurun_out/bench_reference_n$N.json | cut -c1-300
fi
