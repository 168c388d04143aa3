#!/bin/sh
# round-2 GPU call f: full tests (Gaussian option, aggregated assign), launch list of every kernel, long EOS runs at N = 65536
out=gpurun_out/r2f; mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
T="python scripts/profile_target.py --burn 300 --sweeps 5 --all"
$T > $out/target_all.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_all.csv $T > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 $out/target_all.log
for phi in 0.70 0.716; do
  python scripts/eos_run.py --n 65536 --phi $phi --seeds 4 --burn 1000000 --samples 100 --stride 2000 --out $out/eos_long_n64k_phi$phi.json > $out/eos_long_$phi.log 2>&1; echo "eos long $phi rc=$?"; tail -c 700 $out/eos_long_$phi.log
done
