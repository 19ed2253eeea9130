#!/bin/bash
# Round 2, call S: ncu --set full of the literal scan (current build), 4-row variant.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
RJ_EM_DEPTH4=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_emit -s 2 -c 1 -o gpurun_out/r2s_literal -f python scripts/emit_probe.py literal > gpurun_out/r2s_ncu_literal.log 2>&1
tail -2 gpurun_out/r2s_ncu_literal.log
python scripts/ncu_summary.py gpurun_out/r2s_literal.ncu-rep 200000000 | tee gpurun_out/r2s_literal_summary.txt | head -60
