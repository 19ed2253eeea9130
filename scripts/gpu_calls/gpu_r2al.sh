#!/bin/bash
# Round 2, call AL: staging chunk size x threads for pageable uploads (samples/e2e_dropin).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
python - <<'PY' | tee gpurun_out/r2al_staging.txt
import os, subprocess, sys, json
sys.path.insert(0, os.getcwd())
from rejit_b200 import workloads as W
os.makedirs("samples/_build", exist_ok=True)
subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", "include", "samples/e2e_dropin.cc", "-o", "samples/_build/e2e_dropin", "-L", "rejit_b200", "-lrejit_b200", "-Wl,-rpath," + os.path.join(os.getcwd(), "rejit_b200")])
W.fasta_sequence(5_000_000).tofile("samples/_build/seq50.bin")
print("nproc", os.cpu_count())
for kb in (256, 1024, 4096, 8192, 16384):
    for w in (2, 4, 6, 8):
        out = subprocess.run(["samples/_build/e2e_dropin", "samples/_build/seq50.bin", "5"], capture_output=True, text=True,
                             env=dict(os.environ, RJ_STAGE_WIDTH=str(w), RJ_STAGE_CHUNK_KB=str(kb)))
        try:
            d = json.loads(out.stdout.strip().split("\n")[-1])
            print("chunk %5d KB  threads %d  uploaded %.1f GB/s  ms/step %.2f" % (kb, w, d["uploaded_gbs"], d["ms_per_step"]), flush=True)
        except Exception as e:
            print(kb, w, "ERR", out.stderr[-200:], flush=True)
PY
