#!/bin/bash
# Builds an experiment variant of the whole library with extra -D flags:
#   tools/build_variant.sh <name> "<flags>"   ->  tools/build/libb200track_<name>.so   (select with B200TRACK_LIB=<path>)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/a-lightweight-unsupervised-feature-extractor-_b200/csrc
name=$1; flags=$2
mkdir -p $ROOT/tools/build/$name
for f in $CSRC/*.cu; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include $flags \
    -diag-suppress 177 -c $f -o $ROOT/tools/build/$name/$b.o &
done
wait
nvcc -shared -o $ROOT/tools/build/libb200track_$name.so $ROOT/tools/build/$name/*.o -gencode arch=compute_100a,code=sm_100a -cudart shared
rm -rf $ROOT/tools/build/$name
echo built tools/build/libb200track_$name.so
