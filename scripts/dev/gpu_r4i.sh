#!/bin/sh
# final kernel (v18): GPU tests, smoke, all four single-GPU bench lines, --set full of one whole-sweep launch, launch list
out=gpurun_out/r4i; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $out/bench_n16m.json 2> $out/bench_n16m.err; echo "bench rc=$?"
python bench.py --workload n1m_phi0.70 --no-cpu-baseline > $out/bench_n1m.json 2> $out/bench_n1m.err
python bench.py --workload n4m_phi0.30 --no-cpu-baseline > $out/bench_n4m.json 2> $out/bench_n4m.err
python bench.py --workload n16m_phi0.716 --no-cpu-baseline > $out/bench_n16m_716.json 2> $out/bench_n16m_716.err
for f in n16m n1m n4m n16m_716; do python -c "
import json; d=json.load(open('$out/bench_$f.json')); print('$f', '%.4e'%d['value'], 'e2e %.4e'%d['e2e']['value'], 'frac %.3f'%d['roofline']['frac'], d['invariants']['min_d2'], d['invariants']['state_hash'], d['status'])"; done
T="python scripts/profile_target.py --burn 300 --sweeps 5 --bands 1"
ncu --set full --clock-control none --import-source on -k regex:sweep4_kernel -s 300 -c 1 -o $out/sweep4 $T > $out/ncu_sweep4.log 2>&1
echo "sweep4 full rc=$?"; tail -2 $out/ncu_sweep4.log
T2="python scripts/profile_target.py --burn 20 --sweeps 5 --all"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches.csv $T2 > $out/ncu_launches.log 2>&1
echo "launches rc=$?"
