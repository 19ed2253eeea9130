#!/bin/bash
# Round 2, call X: the evidence kept under profiles/: launch list of bench.py, one ncu --set full launch per kernel
# with its summary (scripts/ncu_summary.py), SASS census.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== launch list of bench.py (headline, 2 steps)"
RJ_BENCH_CONFIGS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2x_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2x_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r2x_bench_under_ncu.log | cut -c1-200
cap() {   # name kernel-regex algorithmic-bytes command...
  name=$1; k=$2; alg=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2x_$name -f "$@" > gpurun_out/r2x_ncu_$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/r2x_$name.ncu-rep $alg > gpurun_out/r2x_ncu_${name}_summary.txt 2>&1
  head -22 gpurun_out/r2x_ncu_${name}_summary.txt | sed -n '2,3p;11,12p;20,22p'
}
cap kmer50   k_set_kmer  50335056  python scripts/kmer_probe.py
cap kmer625  k_set_kmer  629186560 python scripts/emit_probe.py kmer625
cap literal  k_scan_emit 200000000 python scripts/emit_probe.py literal
cap c3hits   k_scan_emit 200320000 python scripts/emit_probe.py c3hits
cap c4       k_scan_emit 203130000 python scripts/emit_probe.py c4
cap hat      k_scan_emit 125000000 python scripts/emit_probe.py hat
cap strip    k_scan_emit 128330000 python scripts/emit_probe.py strip
RJ_AB_REPS=1 cap trcount k_translate_count2 625000000 python scripts/ab_run.py iub
RJ_AB_REPS=1 cap trwrite k_translate_write2 1459998180 python scripts/ab_run.py iub
RJ_AB_REPS=1 cap repstage k_replace_stage 1008333411 python scripts/ab_run.py striprep
# the reports themselves are large (the summaries are what is kept): only two travel back
for f in gpurun_out/r2x_*.ncu-rep; do case $f in *kmer50*|*c4.ncu-rep) ;; *) rm -f $f ;; esac; done
ls -la gpurun_out/r2x_*summary.txt | wc -l
