#!/bin/sh
# round-2 GPU call g: Lennard-Jones mode tests + everything else
out=gpurun_out/r2g; mkdir -p $out
python -m pytest tests/test_lj_gpu.py -m gpu -q -x > $out/pytest_lj.log 2>&1; echo "pytest lj rc=$?"; tail -25 $out/pytest_lj.log
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
