#!/bin/bash
# Round 2, call K: streaming loads without L1 allocation (A/B), per-kernel times of the translate passes.
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for v in base ldg; do
  echo "== variant $v"
  RJ_LIB=$PWD/rejit_b200/_variants/lib_$v.so timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b hat strip 2>&1 | tail -8 | tee gpurun_out/r2k_ab_$v.txt
done
echo "== iub launches"
RJ_AB_REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2k_iub_launches.csv python scripts/ab_run.py iub > gpurun_out/r2k_iub.log 2>&1
tail -2 gpurun_out/r2k_iub.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2k_iub_launches.csv")) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iid = hdr.index("ID")
agg = {}
for r in rows[1:]:
    agg.setdefault((r[iid], r[ik][:60]), {})[r[im]] = r[iv]
for (i, k), m in list(agg.items())[-14:]:
    print(i, k, m)
PY
