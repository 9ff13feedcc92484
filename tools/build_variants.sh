#!/bin/bash
# Builds experiment variants of libb200track.so that differ in -D flags of ONE source file of csrc/:
#   tools/build_variants.sh roi_align name1 "-DFLAG=.." name2 "-DFLAG=.." ...   ->  tools/build/libb200track_<name>.so
# Select one at run time with B200TRACK_LIB=tools/build/libb200track_<name>.so (see _lib.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/a-lightweight-unsupervised-feature-extractor-_b200/csrc
make -C $CSRC -j8 > /dev/null
mkdir -p $ROOT/tools/build
SRC=$1; shift
OTHERS=$(ls $CSRC/build/*.o | grep -v "/$SRC.o")
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include $flags \
      -c $CSRC/$SRC.cu -o $ROOT/tools/build/${SRC}_$name.o &&
    nvcc -shared -o $ROOT/tools/build/libb200track_$name.so $OTHERS $ROOT/tools/build/${SRC}_$name.o \
      -gencode arch=compute_100a,code=sm_100a -cudart shared &&
    echo built $name ) &
done
wait
