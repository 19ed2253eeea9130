#!/bin/bash
# Round 2, call J: A/B of k_scan_emit variants (look-back wait, NFA tables in shared memory, tile size), the new
# translate kernels (tests + 625 MB timing).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== replace tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "replace" 2>&1 | tail -8 | tee gpurun_out/r2j_pytest.log
echo "== iub 625 MB"
timeout 600 python scripts/ab_run.py iub 2>&1 | tail -3 | tee gpurun_out/r2j_iub.txt
for v in base spin nosm spin_nosm sleep20 rows96; do
  echo "== variant $v"
  RJ_LIB=$PWD/rejit_b200/_variants/lib_$v.so timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b hat strip 2>&1 | tail -8 | tee gpurun_out/r2j_ab_$v.txt
done
