#!/bin/sh
# usage: gpu_prof2.sh <tag> [launch index]  -- ncu --set full of ONE whole-sweep launch (bands = 1) at N = 2^24
out=gpurun_out/$1; mkdir -p $out; s=${2:-300}
T="python scripts/profile_target.py --burn 300 --sweeps 5 --bands 1"
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s $s -c 1 -o $out/sweep4 $T > $out/ncu_sweep4.log 2>&1
echo "sweep4 full rc=$?"; tail -2 $out/ncu_sweep4.log
