#!/bin/bash
# Round 2, call E: ticket fix of the scan+emit kernels, the new bench.py (all configurations), size sweep of k_set_kmer.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q --timeout 1000 -k "single_pass or fused_pattern_set or kmer_set_long or replace or golden or empty_and_tiny or edges or finish_in_kernel" 2>&1 | tail -25 | tee gpurun_out/r2e_pytest_new.log
echo "== extra"; RJ_EXTRA_REPS=5 RJ_EXTRA_CHAIN_LINES=2000000 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2e_extra.err | tee gpurun_out/r2e_bench_extra.jsonl | cut -c1-330
tail -5 gpurun_out/r2e_extra.err
echo "== kmer sweep"; timeout 600 python scripts/kmer_sweep.py 2>&1 | tee gpurun_out/r2e_kmer_sweep.jsonl
echo "== bench (ours)"; /usr/bin/time -v timeout 1200 python bench.py --steps 10 --warmup 3 2> gpurun_out/r2e_bench.err | tee gpurun_out/r2e_bench_ours.json | cut -c1-600
tail -25 gpurun_out/r2e_bench.err | cut -c1-300
echo "== bench (reference)"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2e_bench_ref.err | tee gpurun_out/r2e_bench_reference.json | cut -c1-400
