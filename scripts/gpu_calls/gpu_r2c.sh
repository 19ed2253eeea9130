#!/bin/bash
# Round 2, call C: fixes of call B + the ReplaceAll kernels; whole GPU tier, timings, the k-mer probe build.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== new tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "single_pass or replace or fused_pattern_set or kmer_set_long" 2>&1 | tail -25 | tee gpurun_out/r2c_pytest_new.log
echo "== whole gpu tier"
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 2>&1 | tail -30 | tee gpurun_out/r2c_pytest.log
echo "== extra"; RJ_EXTRA_REPS=5 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2c_extra.err | tee gpurun_out/r2c_bench_extra.jsonl | cut -c1-330
tail -5 gpurun_out/r2c_extra.err
echo "== launches of the extra cases (100 MB)"
RJ_EXTRA_REPS=1 RJ_EXTRA_BYTES=100000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_extra.csv python scripts/bench_extra.py > gpurun_out/r2c_extra_ncu.log 2>&1
tail -2 gpurun_out/r2c_extra_ncu.log
echo "== kmer probe build"
RJ_NVCC_EXTRA=-DRJ_KMER_PROBE python -m rejit_b200.build --force > gpurun_out/r2c_probe_build.log 2>&1
timeout 300 python scripts/kmer_probe.py 2>&1 | grep -E "probe|bytes" | tail -12 | tee gpurun_out/r2c_kmer_probe.txt | cut -c1-600
python -m rejit_b200.build --force > /dev/null 2>&1
