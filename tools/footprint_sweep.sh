#!/bin/bash
# 64 c2 streams: launch schedule x total CTAs of the front / update kernels x shared-memory split vs the overlapped step.
# Record of an experiment (profiles/r02_footprint_sweep.txt): B200TRACK_DEFAULT_CARVEOUT toggled a source patch that set
# cudaFuncAttributePreferredSharedMemoryCarveout = max on every chain kernel (not kept: slower); fw2 / fw1 are
# tools/build_variant.sh builds with -DB200_TRK_FRONT_WARPS=2 / 1.
run() { echo "== $*"; env "$@" MODES=assoc_only,overlap_prio python tools/group_probe.py 64 nchw c2 2>&1 | tail -1; }
run B200TRACK_CHAIN=3
run B200TRACK_CHAIN=3 B200TRACK_DEFAULT_CARVEOUT=1
run B200TRACK_CHAIN=6
run B200TRACK_CHAIN=3 B200TRACK_FRONT_CTAS=148
run B200TRACK_CHAIN=3 B200TRACK_FRONT_CTAS=148 B200TRACK_UPD_CTAS=148
run B200TRACK_CHAIN=3 B200TRACK_FRONT_CTAS=128 B200TRACK_UPD_CTAS=74
run B200TRACK_CHAIN=3 B200TRACK_FRONT_CTAS=64 B200TRACK_UPD_CTAS=74
run B200TRACK_CHAIN=3 B200TRACK_LIB=$PWD/tools/build/libb200track_fw2.so B200TRACK_FRONT_CTAS=148 B200TRACK_UPD_CTAS=148
run B200TRACK_CHAIN=3 B200TRACK_LIB=$PWD/tools/build/libb200track_fw2.so B200TRACK_FRONT_CTAS=296 B200TRACK_UPD_CTAS=148
run B200TRACK_CHAIN=3 B200TRACK_LIB=$PWD/tools/build/libb200track_fw1.so B200TRACK_FRONT_CTAS=296 B200TRACK_UPD_CTAS=148
run B200TRACK_CHAIN=3 B200TRACK_LIB=$PWD/tools/build/libb200track_fw1.so B200TRACK_FRONT_CTAS=148 B200TRACK_UPD_CTAS=74
