#!/bin/bash
# Build kernel-experiment variants of libb2f_cuda.so into build/variants/<name>.so
# usage: tools/build_variants.sh name "-DFLAG=1 -DOTHER=2" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  d=build/variants/$name; mkdir -p $d
  for f in api costvol warp criterions conv train conv_tc costvol_tc wgrad_tc; do
    nvcc -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $flags -c back2future_b200/csrc/$f.cu -o $d/$f.o 2> $d/$f.ptxas.log &
  done
  wait
  nvcc $ARCH -shared -o build/variants/$name.so $d/api.o $d/costvol.o $d/warp.o $d/criterions.o $d/conv.o $d/train.o $d/conv_tc.o $d/costvol_tc.o $d/wgrad_tc.o -Xlinker --version-script=back2future_b200/csrc/exports.map
  echo built build/variants/$name.so
done
