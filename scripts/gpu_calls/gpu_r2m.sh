#!/bin/bash
# Round 2, call M: L2 bulk prefetch ahead of the streaming loads (A/B).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== main"; timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b hat strip 2>&1 | tail -8 | tee gpurun_out/r2m_ab_main.txt
for v in pf8 pf16 pf32; do
  echo "== variant $v"
  RJ_LIB=$PWD/rejit_b200/_variants/lib_$v.so timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b hat strip 2>&1 | tail -8 | tee gpurun_out/r2m_ab_$v.txt
done
