#!/bin/bash
# Round 2, call R: eight rows in flight in the literal kernels (A/B against four), window evaluation after the last row.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 1000 -k "single_pass or workloads_medium or full_size or golden or reference_test_table or fuzz or long_literals or edges" 2>&1 | tail -8 | tee gpurun_out/r2r_pytest.log
echo "== depth 8 (default)"; timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b 2>&1 | tail -6 | tee gpurun_out/r2r_ab_d8.txt
echo "== depth 4"; RJ_EM_DEPTH4=1 timeout 600 python scripts/ab_run.py lit c4 b 2>&1 | tail -4 | tee gpurun_out/r2r_ab_d4.txt
