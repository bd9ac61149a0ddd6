#!/bin/bash
# build an ablation variant of the extension: tools/build_variant.sh NAME [-DMACRO=VALUE ...]  ->  build/variants/libipp_NAME.so
# (select it at run time with IPP_B200_LIB=build/variants/libipp_NAME.so)
set -e
name=$1; shift
cd "$(dirname "$0")/../ipp_rl_b200/csrc"
mkdir -p ../../build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -ldl -diag-suppress 128 "$@" \
  -o ../../build/variants/libipp_$name.so ipp_engine.cu mcts.cu grf.cu observe.cu experience.cu kalman_blocks.cu fields.cu
echo "built build/variants/libipp_$name.so"
