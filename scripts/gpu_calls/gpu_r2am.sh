#!/bin/bash
# Round 2, call AM: line context known from the start filter (no re-read of the byte before a line start).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --timeout 1000 -k "golden or reference_test_table or fuzz or single_pass or edges or empty_and_tiny or rich_dialect" 2>&1 | tail -4 | tee gpurun_out/r2am_pytest.log
timeout 600 python scripts/ab_run.py hat strip b 2>&1 | tail -3 | tee gpurun_out/r2am_ab.txt
