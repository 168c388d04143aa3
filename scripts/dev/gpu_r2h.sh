#!/bin/sh
# round-2 GPU call h: the reference-function LJ run (fixture), then the LJ tests against it
out=gpurun_out/r2h; mkdir -p $out
./oracle/_ref/ref_harness_lj 16 300 200 > $out/ref_lj_stats.json 2> $out/ref_lj.err; echo "harness rc=$?"; cat $out/ref_lj_stats.json | head -20
cp $out/ref_lj_stats.json tests/golden/ref_lj_stats.json
python -m pytest tests/test_lj_gpu.py tests/test_lj_cpu.py -q > $out/pytest_lj.log 2>&1; echo "pytest lj rc=$?"; tail -15 $out/pytest_lj.log
