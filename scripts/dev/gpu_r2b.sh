#!/bin/sh
# round-2 GPU call b: full GPU test suite, launch list + ncu --set full of every kernel, other workloads
out=gpurun_out/r2b; mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $out/pytest.log
T="python scripts/profile_target.py --burn 300 --sweeps 5 --bands 1"
$T --all > $out/target_all.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_all.csv $T --all > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$T > $out/target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 1 -o $out/sweep4 $T > $out/ncu_sweep4.log 2>&1
echo "sweep4 full rc=$?"
T2="python scripts/profile_target.py --burn 2 --sweeps 2 --bands 1 --all"
$T2 > $out/target2.log 2>&1 && \
ncu --set full --clock-control none -k 'regex:assign|shift_kernel|check_kernel|gr_hist|import4|export4|d2r|sweep_tile' -c 14 -o $out/others $T2 > $out/ncu_others.log 2>&1
echo "others full rc=$?"
for wl in n1m_phi0.70 n16m_phi0.716 n4m_phi0.30; do
  python bench.py --workload $wl --steps 3 --warmup 3 > $out/bench_$wl.json 2> $out/bench_$wl.err; echo "bench $wl rc=$?"
done
ls -la $out
