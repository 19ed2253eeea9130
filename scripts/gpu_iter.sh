#!/bin/bash
# Short iteration run on the GPU box: the set tests, one bench line, the phase
# trace and (RJ_PROFILE=1) an ncu capture of the set scan.  Output in gpurun_out/.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest (set tests)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "${RJ_K:-fused or finish_in_kernel or sample}" --timeout 600 2>&1 | tail -25 | tee gpurun_out/pytest_iter.log
[ "${RJ_BENCH:-1}" = "1" ] && echo "== bench" && timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json | cut -c1-300
tail -3 gpurun_out/bench.err
echo "== phase trace"; RJ_FIN_TRACE=1 timeout 300 python scripts/fin_trace.py 2>&1 | tail -6 | tee gpurun_out/fin_trace.txt | cut -c1-900
[ "${RJ_PHASES:-1}" = "1" ] && echo "== phases" && timeout 300 python scripts/kmer_phases.py 2>&1 | tail -8 | tee gpurun_out/kmer_phases.txt
if [ "${RJ_PROFILE:-0}" = "1" ]; then
echo "== ncu full set scan"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_set_ -s 3 -c 1 -o gpurun_out/prof_set -f python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_set.ncu-rep --page raw --csv > gpurun_out/prof_set_raw.csv 2>/dev/null
fi
echo "== rotate probe"; timeout 300 python scripts/rotate_probe.py 2>&1 | tail -12 | tee gpurun_out/rotate_probe.txt | cut -c1-900
echo "== rotate trace"; RJ_FIN_TRACE=1 timeout 300 python scripts/rotate_probe.py rotate 2>&1 | grep "kmer trace" | tail -3 | tee gpurun_out/rotate_trace.txt | cut -c1-900
