#!/bin/bash
# Round 2, call A: baseline of the round-1 code on this round's box: GPU tests, the
# contract bench line, the other configurations at full size, and a launch list of
# the dense / line-oriented paths.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc > gpurun_out/r2a_nproc.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench_ours.json | cut -c1-400
echo "== extra"; RJ_EXTRA_REPS=5 timeout 900 python scripts/bench_extra.py 2> gpurun_out/r2a_extra.err | tee gpurun_out/r2a_bench_extra.jsonl | cut -c1-400
echo "== launches of the extra cases (100 MB)"
RJ_EXTRA_REPS=1 RJ_EXTRA_BYTES=100000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_extra.csv python scripts/bench_extra.py > gpurun_out/r2a_extra_ncu.log 2>&1
tail -2 gpurun_out/r2a_extra_ncu.log
