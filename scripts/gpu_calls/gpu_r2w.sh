#!/bin/bash
# Round 2, call W (2 GPUs): wider look-back, jrep row, full bench at N=1, headline at N=2.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -8 | tee gpurun_out/r2w_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py lit c3 c3hits c4 b hat strip striprep 2>&1 | tail -9 | tee gpurun_out/r2w_ab.txt
echo "== bench N=1 (all rows)"
timeout 1800 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2w_bench.err | tee gpurun_out/r2w_bench_ours.json | cut -c1-300
tail -5 gpurun_out/r2w_bench.err | cut -c1-300
echo "== bench --gpus 2 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2w_bench2.err | tee gpurun_out/r2w_bench_n2.json | cut -c1-300
tail -5 gpurun_out/r2w_bench2.err | cut -c1-300
