#!/bin/sh
# usage: gpu_tb.sh <tag> : GPU tests (stop at first failure) + bench line
out=gpurun_out/$1; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("value %.4e e2e %.4e ms/sweep %.4f hash %s acc %.6f" % (d["value"], d["e2e"]["value"], d["roofline"]["ms_per_sweep"], d["invariants"]["state_hash"], d["acceptance"]))
PY
