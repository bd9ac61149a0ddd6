#!/bin/bash
# build ablation variants of the library (compile-time switches of step_bulk.cuh) into build/variants/<name>.so
cd "$(dirname "$0")/../ipp_rl_b200/csrc"
SRC="ipp_engine.cu mcts.cu grf.cu observe.cu experience.cu kalman_blocks.cu fields.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -ldl -diag-suppress 128"
build() { name=$1; shift; nvcc $FLAGS "$@" -o ../../build/variants/$name.so $SRC > ../../build/variants/$name.log 2>&1 && echo "built $name" || echo "FAILED $name"; }
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  build $name $(echo $defs | tr ',' ' ') &
done
wait
