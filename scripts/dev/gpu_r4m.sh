#!/bin/sh
# ncu --set full of one whole-sweep launch of the dilute configuration (N = 2^22, phi = 0.30, delta = 0.4: NS = 4 path) and of N = 2^20
out=gpurun_out/r4m; mkdir -p $out
T="python scripts/profile_target.py --n 4194304 --phi 0.30 --delta 0.4 --burn 300 --sweeps 5 --bands 1"
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 1 -o $out/sweep4_dilute $T > $out/ncu_dilute.log 2>&1; echo "dilute rc=$?"
T="python scripts/profile_target.py --n 1048576 --burn 300 --sweeps 5 --bands 1"
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 1 -o $out/sweep4_n1m $T > $out/ncu_n1m.log 2>&1; echo "n1m rc=$?"
