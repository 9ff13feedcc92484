#!/bin/bash
# Builds experiment variants of libb200track.so that differ in -D flags of csrc/roi_align.cu only:
#   tools/build_variants.sh name1 "-DFLAG=.." name2 "-DFLAG=.." ...   ->  tools/build/libb200track_<name>.so
# Select one at run time with B200TRACK_LIB=tools/build/libb200track_<name>.so (see _lib.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/a-lightweight-unsupervised-feature-extractor-_b200/csrc
make -C $CSRC -j8 > /dev/null
mkdir -p $ROOT/tools/build
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include $flags \
      -c $CSRC/roi_align.cu -o $ROOT/tools/build/roi_align_$name.o &&
    nvcc -shared -o $ROOT/tools/build/libb200track_$name.so $CSRC/build/api.o $CSRC/build/assoc_cost.o $CSRC/build/kalman.o \
      $CSRC/build/lsap.o $ROOT/tools/build/roi_align_$name.o $CSRC/build/tracker.o -gencode arch=compute_100a,code=sm_100a -cudart shared &&
    echo built $name ) &
done
wait
