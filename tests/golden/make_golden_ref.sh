#!/bin/sh
# Regenerates tests/golden/ref_kernels_seed*.json: outputs of the REFERENCE's own kernels
# (assign kernel.cu:92-150, V2 shiftCells shiftCells.h:23-112), compiled unmodified from
# /root/reference by oracle/Makefile (target `ref` -> oracle/_ref/ref_harness, harness source
# oracle/ref_harness.cu) and executed on a B200.
#
#   in the build container (has /root/reference, no GPU):   make -C oracle ref
#   on the GPU box (has the prebuilt oracle/_ref, no reference tree):
#       gpurun -- 'sh tests/golden/make_golden_ref.sh gpurun_out/ref'
#   back in the container:  cp gpurun_out/ref/ref_kernels_seed*.json tests/golden/
set -e
out=${1:-gpurun_out/ref}
mkdir -p "$out"
for s in 1 2 3; do ./oracle/_ref/ref_harness $s 40 > "$out/ref_kernels_seed$s.json"; done
./oracle/_ref/ref_harness 7 24 > "$out/ref_kernels_seed7.json"
for s in 4 5 6 8 9 10; do ./oracle/_ref/ref_harness $s 40 > "$out/ref_kernels_seed$s.json"; done
# crowded cells (7-8 particles, immigrant overflow beyond our nmax = 8)
for s in 11 12 13 14; do ./oracle/_ref/ref_harness $s 44 14 > "$out/ref_kernels_seed$s.json"; done
for s in 15 16; do ./oracle/_ref/ref_harness $s 52 18 > "$out/ref_kernels_seed$s.json"; done
# the sub-sweep device functions of subsweep.h on the probes of tests/golden/make_trial_probes.py
./oracle/_ref/ref_harness_v1 < tests/golden/trial_probes_in.txt > "$out/ref_trials.json"
# a short 3-D Lennard-Jones run made of the reference's own device functions / kernels (bug-fixed loop)
./oracle/_ref/ref_harness_lj 16 300 200 > "$out/ref_lj_stats.json"
