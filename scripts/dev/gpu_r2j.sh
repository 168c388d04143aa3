#!/bin/sh
out=gpurun_out/r2j; mkdir -p $out
python scripts/tune_scan.py --n 1048576 --sweeps 4000 > $out/tune_n1m.jsonl 2>&1; cat $out/tune_n1m.jsonl
python scripts/tune_scan.py --n 16777216 --sweeps 600 --bands 4,6,8 --prefetch 148,296,444 > $out/tune_n16m.jsonl 2>&1; cat $out/tune_n16m.jsonl
python scripts/tune_scan.py --n 4194304 --phi 0.30 --delta 0.4 --sweeps 1500 --bands 6,8 --prefetch 296 > $out/tune_n4m.jsonl 2>&1; cat $out/tune_n4m.jsonl
python scripts/tune_scan.py --n 4194304 --phi 0.30 --delta 0.4 --sweeps 1500 --bands 6 --prefetch 296 --extra no_ns4 >> $out/tune_n4m.jsonl 2>&1; tail -1 $out/tune_n4m.jsonl
