#!/bin/bash
# Round 2, call I (2 GPUs): look-back only where needed, rotating texts, fused scan+stitch under torchrun.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -15 | tee gpurun_out/r2i_pytest.log
echo "== extra"; RJ_EXTRA_REPS=5 RJ_EXTRA_CHAIN_LINES=2000000 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2i_extra.err | tee gpurun_out/r2i_bench_extra.jsonl | cut -c1-330
tail -5 gpurun_out/r2i_extra.err
echo "== bench N=1 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2i_bench1.err | tee gpurun_out/r2i_bench_n1.json | cut -c1-1500
tail -5 gpurun_out/r2i_bench1.err | cut -c1-300
echo "== bench --gpus 2 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2i_bench2.err | tee gpurun_out/r2i_bench_n2.json | cut -c1-1500
tail -15 gpurun_out/r2i_bench2.err | cut -c1-300
