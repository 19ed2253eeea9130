#!/bin/bash
# Tuning aid: builds variants of librejit_b200.so with extra -D flags into rejit_b200/_variants/ (git-ignored, travels
# to the GPU box).  usage: ab_build.sh name1 "-DFLAG ..." name2 "-DFLAG ..." ...   (built in parallel)
set -u
cd "$(dirname "$0")/.."
mkdir -p rejit_b200/_variants
C=rejit_b200/csrc
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
      -Xcompiler -Wno-unused-function -Xcompiler -Wno-unknown-pragmas -cudart static -shared $flags \
      -o rejit_b200/_variants/lib_$name.so $C/host/parser.cc $C/host/lower.cc $C/host/automaton.cc $C/host/capi.cc \
      $C/host/regej.cc $C/cuda/engine.cu > rejit_b200/_variants/build_$name.log 2>&1 && echo "built $name" || echo "FAILED $name" ) &
done
wait
