#!/bin/bash
# Round 2, call B: the single-pass scan + emit kernels and k_set_kmer v2 — new tests first (short timeouts: a
# hung kernel must not eat the box), then the whole GPU tier, then timings.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
free -g | head -2 > gpurun_out/r2b_mem.txt; nproc >> gpurun_out/r2b_mem.txt
echo "== new tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout 500 -k "single_pass or empty_and_tiny or golden" 2>&1 | tail -25 | tee gpurun_out/r2b_pytest_new.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "kmer_set_long or fused_pattern_set" 2>&1 | tail -25 | tee gpurun_out/r2b_pytest_kmer.log
echo "== whole gpu tier"
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 | tee gpurun_out/r2b_pytest.log
echo "== extra"; RJ_EXTRA_REPS=5 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2b_extra.err | tee gpurun_out/r2b_bench_extra.jsonl | cut -c1-330
tail -5 gpurun_out/r2b_extra.err
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/r2b_bench.err | tee gpurun_out/r2b_bench_ours.json | cut -c1-400
tail -3 gpurun_out/r2b_bench.err
