#!/bin/sh
# session-3 call e: final kernel (v15): bench lines of the secondary configs, --set full of two consecutive whole-sweep launches, launch list
out=gpurun_out/r4e; mkdir -p $out
python bench.py --workload n1m_phi0.70 --no-cpu-baseline > $out/bench_n1m.json 2> $out/bench_n1m.err; echo "bench n1m rc=$?"
python bench.py --workload n4m_phi0.30 --no-cpu-baseline > $out/bench_n4m.json 2> $out/bench_n4m.err; echo "bench n4m rc=$?"
python bench.py --workload n16m_phi0.716 --no-cpu-baseline > $out/bench_n16m_716.json 2> $out/bench_n16m_716.err; echo "bench n16m 0.716 rc=$?"
T="python scripts/profile_target.py --burn 300 --sweeps 5 --bands 1"
$T > $out/target.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 2 -o $out/sweep4 $T > $out/ncu_sweep4.log 2>&1
echo "sweep4 full rc=$?"; tail -3 $out/ncu_sweep4.log
T2="python scripts/profile_target.py --burn 20 --sweeps 5 --all"
$T2 > $out/target2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches.csv $T2 > $out/ncu_launches.log 2>&1
echo "launches rc=$?"
for f in n1m n4m n16m_716; do python -c "
import json; d=json.load(open('$out/bench_$f.json')); print('$f', '%.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['invariants']['min_d2'], d['status'])"; done
