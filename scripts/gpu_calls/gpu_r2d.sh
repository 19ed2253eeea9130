#!/bin/bash
# Round 2, call D: warp-independent scan+emit, cheaper k_set_kmer flushes.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== new tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "single_pass or fused_pattern_set or kmer_set_long or golden or empty_and_tiny or edges" 2>&1 | tail -25 | tee gpurun_out/r2d_pytest_new.log
echo "== whole gpu tier"
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -30 | tee gpurun_out/r2d_pytest.log
echo "== extra"; RJ_EXTRA_REPS=5 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2d_extra.err | tee gpurun_out/r2d_bench_extra.jsonl | cut -c1-330
tail -5 gpurun_out/r2d_extra.err
