#!/bin/sh
# usage: analyze_prof.sh <dir with sweep4.ncu-rep>   (the library must be the build that was profiled)
d=$1
ncu -i $d/sweep4.ncu-rep --page source --csv --print-source sass > $d/sass.csv 2>/dev/null
tmp=$(mktemp -d); (cd $tmp && cuobjdump -xelf pmc_sweep4 /root/repo/parallel-monte-carlo_b200/libpmc_b200.so >/dev/null)
nvdisasm -g -c $tmp/pmc_sweep4.sm_100a.cubin > $d/sweep4.asm
python scripts/ncu_by_region.py $d/sass.csv $d/sweep4.asm "sweep4_kernelILi4ELb1" > $d/by_region.txt
python scripts/ncu_summary.py $d/sweep4.ncu-rep > $d/summary.txt 2>/dev/null
cat $d/summary.txt; cat $d/by_region.txt
