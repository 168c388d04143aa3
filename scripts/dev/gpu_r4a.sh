#!/bin/sh
# round-2 session-3 call a: state of the tree (v13): GPU tests, bench lines, launch list, --set full of one whole-sweep launch
out=gpurun_out/r4a; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/pytest.log
python bench.py > $out/bench_n16m.json 2> $out/bench_n16m.err; echo "bench rc=$?"; cat $out/bench_n16m.json
python bench.py --workload n1m_phi0.70 --no-cpu-baseline > $out/bench_n1m.json 2> $out/bench_n1m.err; echo "bench n1m rc=$?"; cat $out/bench_n1m.json
python bench.py --workload n4m_phi0.30 --no-cpu-baseline > $out/bench_n4m.json 2> $out/bench_n4m.err; echo "bench n4m rc=$?"; cat $out/bench_n4m.json
T="python scripts/profile_target.py --burn 300 --sweeps 5 --bands 1"
$T > $out/target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 1 -o $out/sweep4 $T > $out/ncu_sweep4.log 2>&1
echo "sweep4 full rc=$?"; tail -3 $out/ncu_sweep4.log
T2="python scripts/profile_target.py --burn 20 --sweeps 5 --all"
$T2 > $out/target2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches.csv $T2 > $out/ncu_launches.log 2>&1
echo "launches rc=$?"
