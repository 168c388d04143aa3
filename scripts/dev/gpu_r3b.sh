#!/bin/sh
out=gpurun_out/r3b; mkdir -p $out
PMC_SWEEPS=300 timeout 300 python scripts/dev/quick16m.py > $out/quick.log 2>&1; cat $out/quick.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json
