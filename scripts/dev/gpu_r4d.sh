#!/bin/sh
# session-3 call d: full GPU tests + bench line after the new trial-displacement construction
out=gpurun_out/r4d; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $out/bench_n16m.json 2> $out/bench_n16m.err; echo "bench rc=$?"; tail -3 $out/bench_n16m.err
python -c "
import json; d=json.load(open('$out/bench_n16m.json')); print('%.4e'%d['value'], d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'], d['acceptance'], d['invariants']['state_hash'], d['roofline']['frac'], d['cpu_baseline']['value'])"
