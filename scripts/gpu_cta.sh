#!/bin/bash
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest (set tests)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or finish_in_kernel or sample" --timeout 600 2>&1 | tail -5
RJ_FIN_TRACE=1 RJ_FIN_TRACE_CTA=0,29,60,100,147 timeout 300 python scripts/fin_trace.py 2>&1 | grep "kmer" | tail -6 | tee gpurun_out/cta_trace.txt | cut -c1-1200
echo "== phases"; timeout 300 python scripts/kmer_phases.py 2>&1 | tail -8 | tee gpurun_out/kmer_phases.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json | cut -c1-200
