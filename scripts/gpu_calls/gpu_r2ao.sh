#!/bin/bash
# Round 2, call AO (the round's last 78 GPU-seconds): the fused-rebuild test in its final form (call AN's version held
# two `a.*` cases over a 1 MB line: 490 k speculative starts x 490 kB each, six minutes of NFA runs).
set -u
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q --timeout 55 --durations=3 -k "replace_all_fused" 2>&1 | tail -12 | tee gpurun_out/r2ao_pytest.log
