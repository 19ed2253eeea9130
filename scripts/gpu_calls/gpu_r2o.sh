#!/bin/bash
# Round 2, call O: replace index kernel, translate pre-filter, tile size A/B.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 1000 -k "replace or regexdna or single_pass" 2>&1 | tail -8 | tee gpurun_out/r2o_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py iub striprep lit c3 c4 b 2>&1 | tail -8 | tee gpurun_out/r2o_ab.txt
for v in rows32 rows48; do
  echo "== variant $v"
  RJ_LIB=$PWD/rejit_b200/_variants/lib_$v.so timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b hat strip 2>&1 | tail -8 | tee gpurun_out/r2o_ab_$v.txt
done
