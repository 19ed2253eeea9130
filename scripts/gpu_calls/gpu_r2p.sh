#!/bin/bash
# Round 2, call P: tile size sweep at run time, memory-side ceiling of the tiling (filter replaced by XORs).
set -u
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
for rws in 16 24 32 48 64; do
  echo "== tile rows $rws"
  RJ_EM_TILE_ROWS=$rws timeout 600 python scripts/ab_run.py lit c3hits c4 b hat strip 2>&1 | tail -7 | tee gpurun_out/r2p_rows_$rws.txt
done
for v in ro ro32; do
  echo "== variant $v"
  RJ_LIB=$PWD/rejit_b200/_variants/lib_$v.so timeout 600 python scripts/ab_run.py lit 2>&1 | tail -2 | tee gpurun_out/r2p_ab_$v.txt
done
echo "== defaults"; timeout 600 python scripts/ab_run.py lit c3 c3hits c4 b hat strip striprep iub 2>&1 | tail -10 | tee gpurun_out/r2p_ab.txt
