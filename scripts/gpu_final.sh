#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest gpu (quick)"; timeout 300 python -m pytest tests -m gpu -q -x --timeout 250 -k "not full_size" 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_quick.log
echo "== lit probe"; timeout 120 python scripts/lit_probe.py 2>&1 | tee gpurun_out/lit_probe.jsonl | cut -c1-300
echo "== pytest gpu (full size)"; timeout 200 python -m pytest tests -m gpu -q -x --timeout 190 -k "full_size" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_full.log
echo "== bench"; timeout 120 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json | cut -c1-200
