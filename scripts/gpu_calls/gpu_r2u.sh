#!/bin/bash
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python scripts/size_sweep.py 2>&1 | tail -5 | tee gpurun_out/r2u_size_sweep.txt
RJ_EM_DEPTH4=1 timeout 600 python scripts/size_sweep.py 2>&1 | tail -5 | tee gpurun_out/r2u_size_sweep_d4.txt
