#!/bin/bash
# Builds tools/build/libb200track_timing.so: the whole library with -DB200_TRK_TIMING (global-timer spans and phase stamps;
# tools/timeline_probe.py, tools/fused_timeline.py, tools/trk_timing.py).  Select it with B200TRACK_LIB=<that path>.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC=$ROOT/a-lightweight-unsupervised-feature-extractor-_b200/csrc
mkdir -p $ROOT/tools/build/timing
for f in $CSRC/*.cu; do
  b=$(basename $f .cu)
  if [ ! -f $ROOT/tools/build/timing/$b.o ] || [ -n "$(find $CSRC $ROOT/include -newer $ROOT/tools/build/timing/$b.o -name '*.cu*' -o -newer $ROOT/tools/build/timing/$b.o -name '*.h')" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I$ROOT/include -DB200_TRK_TIMING \
      -diag-suppress 177 -c $f -o $ROOT/tools/build/timing/$b.o &
  fi
done
wait
nvcc -shared -o $ROOT/tools/build/libb200track_timing.so $ROOT/tools/build/timing/*.o -gencode arch=compute_100a,code=sm_100a -cudart shared
echo built tools/build/libb200track_timing.so
