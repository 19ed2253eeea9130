#!/bin/bash
# Round 2, call L (2 GPUs): k_set_kmer with the top-CTA report, two-phase translate write, scan+emit defaults.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -15 | tee gpurun_out/r2l_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py kmer50 kmer625 iub lit c3 c3hits c4 b hat strip 2>&1 | tail -12 | tee gpurun_out/r2l_ab.txt
echo "== bench N=1 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2l_bench1.err | tee gpurun_out/r2l_bench_n1.json | cut -c1-300
tail -5 gpurun_out/r2l_bench1.err | cut -c1-300
echo "== bench --gpus 2 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r2l_bench2.err | tee gpurun_out/r2l_bench_n2.json | cut -c1-300
tail -8 gpurun_out/r2l_bench2.err | cut -c1-300
