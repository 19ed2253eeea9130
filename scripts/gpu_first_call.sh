#!/bin/bash
# First gpurun call of a round: everything written since the last GPU run gets its first execution here.
#   gpurun --timeout 1500 -- bash scripts/gpu_first_call.sh
# GPU tier WITHOUT -x (one failing new test must not hide the others), durations, then the samples with
# their phase traces, then one bench line.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --durations=12 -rfEs 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== jrep on a synthetic tree (512 MB, 8192 files)"
python - <<'PY'
import os, sys
sys.path.insert(0, ".")
from rejit_b200 import workloads as W
blob = W.source_blob(64 << 20, seed=9).tobytes()
for rep in range(8):
    for i in range(1024):
        d = "/tmp/jrep_big/r%d/d%02d" % (rep, i % 32)
        os.makedirs(d, exist_ok=True)
        open(os.path.join(d, "f%04d.c" % i), "wb").write(blob[i * 65536:(i + 1) * 65536])
PY
g++ -std=c++11 -O2 -Iinclude samples/jrep.cc -Lrejit_b200 -lrejit_b200 -Wl,-rpath,$PWD/rejit_b200 -lpthread -o /tmp/jrep
for args in "" "-j0" "--batch-bytes=268435456" "--batch-bytes=16777216"; do
  for pat in 'qqqzzzqqq' ';
}'; do
    ( time JREP_TRACE=1 /tmp/jrep $args -r -n "$pat" /tmp/jrep_big | wc -c ) 2>&1 | grep -v "^$\|user\|sys" | tr '\n' ' '; echo " [args '$args' pattern $(printf %q "$pat")]"
  done
done | tee gpurun_out/jrep_trace.txt
echo "== bench_engine (reference harness table)"
g++ -std=c++11 -O2 -Iinclude samples/bench_engine.cc -Lrejit_b200 -lrejit_b200 -Wl,-rpath,$PWD/rejit_b200 -o /tmp/bench_engine
for extra in "" "--resident=1"; do
  /tmp/bench_engine '([complex]|(regexp)){2,7}abcdefgh(at|the|[e-nd]as well)' --iterations=20 --low_char=0 --high_char=z \
      --size=4096,65536,1048576,16777216,268435456 $extra
done 2>&1 | tee gpurun_out/bench_engine.txt
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_ours.json | cut -c1-300
tail -3 gpurun_out/bench.err
ls -la gpurun_out
