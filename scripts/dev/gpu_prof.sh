#!/bin/sh
# usage: gpu_prof.sh <tag>   -- ncu --set full of one whole-sweep launch (bands = 1) at N = 2^24 after 300 burn-in sweeps
out=gpurun_out/$1; mkdir -p $out
T="python scripts/profile_target.py --burn 300 --sweeps 5 --bands 1"
$T > $out/target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 1 -o $out/sweep4 $T > $out/ncu_sweep4.log 2>&1
echo "sweep4 full rc=$?"; tail -3 $out/ncu_sweep4.log
