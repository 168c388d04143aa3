#!/bin/sh
# per-sweep instruction counts and durations of 12 consecutive whole-sweep launches (x- and y-shift sweeps)
out=gpurun_out/r4f; mkdir -p $out
T="python scripts/profile_target.py --burn 298 --sweeps 14 --bands 1"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__grid_size --clock-control none -k regex:sweep4_kernel -s 298 -c 14 --csv --log-file $out/per_sweep.csv $T > $out/ncu.log 2>&1
echo rc=$?
