#!/bin/bash
# ROI Align with L2 evict-first hints on its output stores (l2h1) / stores and footprint loads (l2h2) vs none:
# ROI alone, chain alone, overlapped step (variants built by tools/build_variant.sh <name> "-DB200_ROI_L2_HINT=1|2").
for lib in "" l2h1 l2h2; do
  echo "== lib=${lib:-product}"
  if [ -n "$lib" ]; then export B200TRACK_LIB=$PWD/tools/build/libb200track_$lib.so; else unset B200TRACK_LIB; fi
  for cfg in "64 nchw c2" "64 nchw c5" "8 nchw c5"; do
    MODES=roi_only,assoc_only,overlap_prio python tools/group_probe.py $cfg 2>&1 | tail -1
  done
done
