#!/bin/bash
# Stream-group step for each association-chain mode (B200TRACK_CHAIN = 6 | 3 | 2): chain alone and overlapped with ROI Align.
for cfg in "64 nchw c2" "64 nchw c5" "16 nchw c5" "8 nchw c5"; do
  for m in 6 3 2; do
    echo "== B200TRACK_CHAIN=$m  $cfg"
    B200TRACK_CHAIN=$m MODES=assoc_only,overlap_prio python tools/group_probe.py $cfg 2>&1 | tail -1
  done
done
