#!/bin/bash
# Round-2 profiler evidence (run under gpurun, one GPU; every command first exits 0 without ncu):
#   1. launch list of bench.py's timed steps + roofline probe     -> gpurun_out/r02_launches_bench.csv
#   2. launch list of the single-stream (two-launch, cluster) step -> gpurun_out/r02_launches_single.csv
#   3. ncu --set full of the three-launch chain's kernels          -> gpurun_out/chain3_r2.ncu-rep
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
$BENCH > gpurun_out/plain_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv \
    $BENCH > gpurun_out/ncu_bench_short.log 2>&1
python tools/chain_once.py 1 > gpurun_out/plain_chain_single.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/r02_launches_single.csv \
    python tools/chain_once.py 1 > gpurun_out/ncu_chain_single.log 2>&1
python tools/chain_once.py 64 > gpurun_out/plain_chain64.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"front_kernel|back_kernel|update_kernel" -s 105 -c 3 \
    -f -o gpurun_out/chain3_r2 python tools/chain_once.py 64 > gpurun_out/ncu_chain3.log 2>&1
ls -la gpurun_out/*.csv gpurun_out/chain3_r2.ncu-rep
