#!/bin/bash
# Runs tools/roi_bench.py (NCHW, b200 only) for the product library and every tools/build/libb200track_*.so variant.
cases=${1:-g64}
echo "== product"; python tools/roi_bench.py $cases --b200 --nchw 2>&1 | grep -v "nhwc" 
for f in tools/build/libb200track_*.so; do
  echo "== $f"; B200TRACK_LIB=$PWD/$f python tools/roi_bench.py $cases --b200 --nchw 2>&1 | grep -v "nhwc"
done
