#!/bin/bash
# Round 2, call N: early ticket + done counter without a round trip, pinned arena for pageable uploads, translate profile.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== tests"
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 1000 2>&1 | tail -15 | tee gpurun_out/r2n_pytest.log
echo "== ab_run"
timeout 900 python scripts/ab_run.py kmer50 kmer625 iub lit c3 c3hits c4 b hat strip 2>&1 | tail -12 | tee gpurun_out/r2n_ab.txt
echo "== e2e_dropin staging sweep"
python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
from rejit_b200 import workloads as W
os.makedirs("samples/_build", exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", "include", "samples/e2e_dropin.cc", "-o", "samples/_build/e2e_dropin", "-L", "rejit_b200", "-lrejit_b200", "-Wl,-rpath," + os.path.join(os.getcwd(), "rejit_b200")])
W.fasta_sequence(5_000_000).tofile("samples/_build/seq50.bin")
for w in (2, 4, 6, 8, 10, 14):
    out = subprocess.run(["samples/_build/e2e_dropin", "samples/_build/seq50.bin", "5"], capture_output=True, text=True, env=dict(os.environ, RJ_STAGE_WIDTH=str(w)))
    print(w, out.stdout.strip()[:170], out.stderr[-200:], flush=True)
PY
echo "== ncu translate write"
RJ_AB_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_translate_write2 -s 1 -c 1 -o gpurun_out/r2n_trwrite -f python scripts/ab_run.py iub > gpurun_out/r2n_ncu_trwrite.log 2>&1
tail -2 gpurun_out/r2n_ncu_trwrite.log
