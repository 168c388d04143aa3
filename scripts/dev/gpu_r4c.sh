#!/bin/sh
# session-3 call c: new trial-displacement construction: reference answers for the regenerated probes, GPU tests, timing
out=gpurun_out/r4c; mkdir -p $out
./oracle/_ref/ref_harness_v1 < tests/golden/trial_probes_in.txt > $out/ref_trials.json; echo "harness rc=$?"
cp $out/ref_trials.json tests/golden/ref_trials.json
timeout 300 python -m pytest tests/test_oracle_cpu.py -q -x --timeout 120 -k "trial or probe" > $out/pytest_cpu.log 2>&1; tail -3 $out/pytest_cpu.log
timeout 1200 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
PMC_SWEEPS=300 python scripts/dev/quick16m.py
PMC_SWEEPS=300 python scripts/dev/quick16m.py
