#!/bin/bash
# Round 2, call AB: two-level look-back.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -8 | tee gpurun_out/r2ab_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py lit c3 c3hits c4 b hat strip striprep c4big litbig 2>&1 | tail -11 | tee gpurun_out/r2ab_ab.txt
