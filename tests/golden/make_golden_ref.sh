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
