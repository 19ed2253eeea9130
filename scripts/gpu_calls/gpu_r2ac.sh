#!/bin/bash
# Round 2, call AC: deferred placement (two candidate lists), generic early evaluation A/B.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -8 | tee gpurun_out/r2ac_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py lit c3 c3hits c4 b hat strip striprep c4big 2>&1 | tail -10 | tee gpurun_out/r2ac_ab.txt
for v in gs256 gs192; do
  echo "== variant $v"
  RJ_LIB=$PWD/rejit_b200/_variants/lib_$v.so timeout 600 python scripts/ab_run.py hat strip 2>&1 | tail -2 | tee gpurun_out/r2ac_ab_$v.txt
done
