#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== probe"; (cd scripts/probe && nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o launch_overhead launch_overhead.cu 2>&1 | tail -2; timeout 60 ./launch_overhead) | tee gpurun_out/launch_overhead.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json | cut -c1-300
tail -3 gpurun_out/bench.err
echo "== finish trace"; RJ_FIN_TRACE=1 timeout 300 python scripts/fin_trace.py 2>&1 | tail -8 | tee gpurun_out/fin_trace.txt | cut -c1-600
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=8 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
