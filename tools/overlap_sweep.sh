#!/bin/bash
# Overlapped 64-stream step (and its parts) for the product library and the variants built by tools/build_variant.sh.
for lib in "" "$@"; do
  echo "== lib=${lib:-product}"
  if [ -n "$lib" ]; then export B200TRACK_LIB=$PWD/tools/build/libb200track_$lib.so; else unset B200TRACK_LIB; fi
  MODES=roi_only,assoc_only,overlap_prio python tools/group_probe.py 64 nchw c2 2>&1 | tail -1
done
