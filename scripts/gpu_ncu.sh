#!/bin/bash
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 1 --warp-sampling-buffer-size 536870912 -k regex:k_set_kmer -s 3 -c 1 -o gpurun_out/prof_kmer -f python scripts/fin_trace.py > gpurun_out/ncu_kmer.log 2>&1
tail -3 gpurun_out/ncu_kmer.log
ls -la gpurun_out/prof_kmer.ncu-rep
