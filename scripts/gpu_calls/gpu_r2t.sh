#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 scripts/probe/read_bw 2>&1 | tee gpurun_out/r2t_read_bw.txt
