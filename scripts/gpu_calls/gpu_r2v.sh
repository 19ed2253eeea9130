#!/bin/bash
# Round 2, call V: everything as it stands: GPU tests, the tuning cases, bench.py with every configuration (N=1).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -8 | tee gpurun_out/r2v_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py kmer50 kmer625 iub striprep lit c3 c3hits c4 b hat strip 2>&1 | tail -12 | tee gpurun_out/r2v_ab.txt
echo "== bench (ours)"; date +%s.%N > gpurun_out/r2v_t0
timeout 1800 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2v_bench.err | tee gpurun_out/r2v_bench_ours.json | cut -c1-600
date +%s.%N > gpurun_out/r2v_t1
tail -12 gpurun_out/r2v_bench.err | cut -c1-300
