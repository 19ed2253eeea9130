#!/bin/bash
# Round 2, call F: no per-tile fence in scan+emit; bench.py with every configuration.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --timeout 1000 -k "single_pass or fused_pattern_set or golden or empty_and_tiny or edges or finish_in_kernel or replace_all_set" 2>&1 | tail -15 | tee gpurun_out/r2f_pytest_new.log
echo "== extra"; RJ_EXTRA_REPS=5 RJ_EXTRA_CHAIN_LINES=2000000 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2f_extra.err | tee gpurun_out/r2f_bench_extra.jsonl | cut -c1-330
tail -5 gpurun_out/r2f_extra.err
echo "== bench (ours)"; date +%s.%N > gpurun_out/r2f_t0; timeout 1500 python bench.py --steps 10 --warmup 3 2> gpurun_out/r2f_bench.err | tee gpurun_out/r2f_bench_ours.json | cut -c1-800; date +%s.%N > gpurun_out/r2f_t1
tail -25 gpurun_out/r2f_bench.err | cut -c1-300
