#!/bin/bash
# Round 2, call Q: read-only bandwidth probe; two-phase window evaluation.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== read_bw"; timeout 300 scripts/probe/read_bw 2>&1 | tee gpurun_out/r2q_read_bw.txt
echo "== tests"; timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q --timeout 1000 -k "single_pass or workloads_medium or full_size or golden or reference_test_table or fuzz" 2>&1 | tail -8 | tee gpurun_out/r2q_pytest.log
echo "== ab"; timeout 600 python scripts/ab_run.py c3 c3hits 2>&1 | tail -3 | tee gpurun_out/r2q_ab.txt
