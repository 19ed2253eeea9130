#!/bin/bash
# Round 2, call AJ: DFA kernels with round-robin sub-regions; fresh ncu summaries of the scan+emit kernels.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q --timeout 800 -k "dfa or fused_pattern_set or kmer or workloads_medium or golden or fixed_length or reference_test_table" 2>&1 | tail -5 | tee gpurun_out/r2aj_pytest.log
echo "== bench N=1 (headline only)"
RJ_BENCH_CONFIGS=0 timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/r2aj_bench1.err | tee gpurun_out/r2aj_bench_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['per_pattern_calls'])"
cap() {
  name=$1; k=$2; alg=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2aj_$name -f "$@" > gpurun_out/r2aj_ncu_$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2aj_$name.ncu-rep $alg > gpurun_out/r2aj_ncu_${name}_summary.txt 2>&1
  sed -n '2,3p;20,22p' gpurun_out/r2aj_ncu_${name}_summary.txt
  rm -f gpurun_out/r2aj_$name.ncu-rep
}
cap literal  k_scan_emit 200000000 python scripts/emit_probe.py literal
cap c3hits   k_scan_emit 200320000 python scripts/emit_probe.py c3hits
cap c4       k_scan_emit 203130000 python scripts/emit_probe.py c4
cap hat      k_scan_emit 125000000 python scripts/emit_probe.py hat
