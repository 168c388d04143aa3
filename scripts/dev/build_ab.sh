#!/bin/sh
# builds libpmc_expA.so from HEAD and libpmc_expB.so from the working tree (for scripts/dev/exp_ab.sh)
set -e
L=parallel-monte-carlo_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared"
rm -rf /tmp/pmc_head && mkdir -p /tmp/pmc_head && git archive ${1:-HEAD} $L/csrc include | tar -x -C /tmp/pmc_head
SRCS="pmc_api.cu pmc_cells.cu pmc_sweep.cu pmc_sweep4.cu pmc_lj.cu"
(cd /tmp/pmc_head/$L/csrc && nvcc $FLAGS -o /root/repo/$L/libpmc_expA.so $SRCS -lcudart -ldl) &
(cd $L/csrc && nvcc $FLAGS -o ../libpmc_expB.so $SRCS -lcudart -ldl) &
wait
ls -la $L/libpmc_exp*.so
