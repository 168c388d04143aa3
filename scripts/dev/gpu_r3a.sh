#!/bin/sh
# round-2 session-2 call a: v11 (merged shift + store pass): GPU tests, quick timing
out=gpurun_out/r3a; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest.log
PMC_SWEEPS=300 timeout 300 python scripts/dev/quick16m.py > $out/quick.log 2>&1; cat $out/quick.log
