#!/bin/bash
# Runs on the GPU box under gpurun: GPU test tier, one bench line, ncu launch
# list + one full capture of the dominant kernel.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -5
echo "== pytest gpu" ; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 1500 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json
tail -5 gpurun_out/bench.err
echo "== bench reference" ; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_reference.json
if [ "${RJ_EXTRA:-0}" = "1" ]; then
echo "== finish trace"; timeout 300 python scripts/fin_trace.py 2>&1 | tail -8 | tee gpurun_out/fin_trace.txt | cut -c1-300
echo "== bench_extra"; timeout 1200 python scripts/bench_extra.py 2>&1 | tee gpurun_out/bench_extra.jsonl | cut -c1-400
echo "== ncu launches of bench_extra"; RJ_EXTRA_REPS=2 RJ_EXTRA_BYTES=500000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extra.csv python scripts/bench_extra.py > gpurun_out/ncu_extra.log 2>&1; tail -2 gpurun_out/ncu_extra.log | cut -c1-200
fi
if [ "${RJ_SWEEP:-0}" = "1" ]; then
for w in 6 10 14 18; do echo "== sweep RJ_DFA_WARPS=$w"; RJ_DFA_WARPS=$w timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"; done
fi
if [ "${RJ_PROFILE:-1}" = "1" ]; then
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "== ncu full set scan"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_set_ -s 3 -c 1 -o gpurun_out/prof_set -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
