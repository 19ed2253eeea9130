#!/bin/bash
# Round 2, call AN: ReplaceAll fused into the single-pass generic scan (k_scan_emit<generic, rebuild>): parity, then
# the strip of a 508 MB FASTA file with and without the fusion.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q --timeout 380 -k "replace_all_fused or replace_all_on_device or regexdna_chain_at_size" 2>&1 | tail -15 | tee gpurun_out/r2an_pytest.log
( timeout 120 python scripts/ab_run.py striprep strip hat 2>&1 | tail -3
  RJ_NO_FUSED_REBUILD=1 timeout 120 python scripts/ab_run.py striprep 2>&1 | tail -1 | sed 's/^/no-fusion: /' ) | tee gpurun_out/r2an_ab.txt
