#!/bin/bash
# Round 2, call H: ncu --set full of the scan kernels (one launch each), staging width sweep of the drop-in path.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for c in literal strip hat c3hits c4 kmer625; do
  k=k_scan_emit; [ $c = kmer625 ] && k=k_set_kmer
  echo "== ncu $c"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2h_$c -f python scripts/emit_probe.py $c > gpurun_out/r2h_ncu_$c.log 2>&1
  tail -2 gpurun_out/r2h_ncu_$c.log
  ncu -i gpurun_out/r2h_$c.ncu-rep --page raw --csv > gpurun_out/r2h_${c}_raw.csv 2>/dev/null
done
echo "== staging sweep (e2e_dropin)"
python - <<'PY'
import os, subprocess, json, sys
sys.path.insert(0, os.getcwd())
from rejit_b200 import workloads as W
os.makedirs("samples/_build", exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", "include", "samples/e2e_dropin.cc", "-o", "samples/_build/e2e_dropin", "-L", "rejit_b200", "-lrejit_b200", "-Wl,-rpath," + os.path.join(os.getcwd(), "rejit_b200")])
W.fasta_sequence(5_000_000).tofile("samples/_build/seq50.bin")
for w in (1, 2, 4, 8, 12, 14):
    out = subprocess.run(["samples/_build/e2e_dropin", "samples/_build/seq50.bin", "5"], capture_output=True, text=True, env=dict(os.environ, RJ_STAGE_WIDTH=str(w)))
    print(w, out.stdout.strip()[:200], out.stderr[-200:], flush=True)
PY
