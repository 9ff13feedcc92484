#!/bin/bash
# Three-launch chain, 64 c2 streams: total CTAs of the front / update kernels vs the overlapped step.
run() { echo "== $*"; env "$@" B200TRACK_CHAIN=3 MODES=assoc_only,overlap_prio python tools/group_probe.py 64 nchw c2 2>&1 | tail -1; }
run X=default
run B200TRACK_FRONT_CTAS=128
run B200TRACK_FRONT_CTAS=192
run B200TRACK_FRONT_CTAS=256
run B200TRACK_FRONT_CTAS=148 B200TRACK_UPD_CTAS=148
run B200TRACK_FRONT_CTAS=256 B200TRACK_UPD_CTAS=148
