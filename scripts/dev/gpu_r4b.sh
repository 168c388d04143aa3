#!/bin/sh
# session-3 call b: non-blocking pmc_run_host (two handles), bench with the pipelined e2e leg
out=gpurun_out/r4b; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "run_host" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
python bench.py --no-cpu-baseline > $out/bench_n16m.json 2> $out/bench_n16m.err; echo "bench rc=$?"; tail -3 $out/bench_n16m.err
python -c "
import json; d=json.load(open('$out/bench_n16m.json')); print('%.4e'%d['value'], d['ms_per_step'], d['e2e'])"
