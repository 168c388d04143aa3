#!/bin/sh
# static SASS size of the fused sweep kernel <4, true> in a library (default: the in-tree build) + opcode histogram head
lib=${1:-parallel-monte-carlo_b200/libpmc_b200.so}
cuobjdump -sass $lib | awk '/Function : .*sweep4_kernelILi4ELb1/{on=1;next} /Function :/{on=0} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | awk '{op=$1; if (op ~ /^@/) op=$2; sub(/\..*/,"",op); c[op]++; n++} END{print "total",n; for(k in c) print c[k],k}' | sort -rn | head -16
