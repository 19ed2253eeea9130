#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python scripts/ab_run.py lit c3 c3hits c4 b hat strip striprep iub 2>&1 | tail -10 | tee gpurun_out/r2y_ab.txt
timeout 600 python scripts/jrep_bench.py 2>&1 | tail -1 | tee gpurun_out/r2y_jrep.json | cut -c1-700
