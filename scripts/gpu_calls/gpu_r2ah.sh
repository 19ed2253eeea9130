#!/bin/bash
# Round 2, call AH: per-warp done count; everything: tests, tuning cases, bench with all rows.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -8 | tee gpurun_out/r2ah_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py kmer50 kmer625 iub lit c3 c3hits c4 b hat strip striprep c4big litbig 2>&1 | tail -14 | tee gpurun_out/r2ah_ab.txt
echo "== bench N=1 (all rows)"
timeout 1800 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2ah_bench.err | tee gpurun_out/r2ah_bench_ours.json | cut -c1-200
tail -5 gpurun_out/r2ah_bench.err | cut -c1-300
