#!/bin/sh
out=gpurun_out/r3i; mkdir -p $out
./oracle/_ref/ref_harness_v1 < tests/golden/trial_probes_in.txt > $out/ref_trials.json; echo "harness rc=$?"
cp $out/ref_trials.json tests/golden/ref_trials.json
sh scripts/dev/gpu_tb.sh r3i
timeout 300 python -m pytest tests/test_oracle_cpu.py -q -x -k "trial or probe" > $out/pytest_cpu.log 2>&1; tail -3 $out/pytest_cpu.log
