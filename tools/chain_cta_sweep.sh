for lib in "" tools/build/libb200track_c1w2.so tools/build/libb200track_c1w1.so; do
 for bt in 256 128 64; do
  echo "== lib=${lib:-product} begin_threads=$bt"
  for cfg in "8 nchw c5" "64 nchw c2"; do
   if [ -n "$lib" ]; then export B200TRACK_LIB=$PWD/$lib; else unset B200TRACK_LIB; fi
   B200_TRK_BEGIN_THREADS=$bt MODES=assoc_only,overlap_prio python tools/group_probe.py $cfg 2>&1 | tail -1
  done
 done
done
