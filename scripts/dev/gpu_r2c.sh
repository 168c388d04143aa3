#!/bin/sh
# round-2 GPU call c: pair-barrier kernel: parity, A/B against the CTA-wide barrier, racecheck
out=gpurun_out/r2c; mkdir -p $out
python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_pairbar.json 2> $out/bench_pairbar.err; echo "bench pair rc=$?"
PMC_LIB_PATH=$PWD/parallel-monte-carlo_b200/libpmc_b200_ctabar.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_ctabar.json 2> $out/bench_ctabar.err; echo "bench cta rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_pairbar2.json 2> $out/bench_pairbar2.err
for f in pairbar ctabar pairbar2; do python -c "
import json; d=json.load(open('$out/bench_$f.json')); print('$f', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'])"; done
python scripts/sanitize_target.py > $out/sanitize_plain.log 2>&1 && \
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_target.py > $out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $out/racecheck.log
