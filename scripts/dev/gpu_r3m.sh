#!/bin/sh
# session-2 multi-GPU call (N GPUs given as $1): slab bit identity + 256M bench line
N=$1; out=gpurun_out/r3m_$N; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
python -m pytest tests/test_multi_gpu.py -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for mode in "262144 7" "4194304 9" "1048576 12 crowded"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 scripts/slab_worker.py $mode 2>&1 | grep SLAB | tee -a $out/slab_worker.log
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus $N --steps 3 --warmup 3 > $out/bench_n256m_${N}gpu.json 2> $out/bench_n256m_${N}gpu.err; echo "bench ${N}gpu rc=$?"
python -c "
import json; d=json.load(open('$out/bench_n256m_${N}gpu.json')); print('${N}gpu', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['invariants']['min_d2'], 'e2e %.4e'%d['e2e']['value'])"
if [ "$N" = "2" ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus 1 --steps 3 --warmup 3 > $out/bench_n256m_1gpu.json 2> $out/bench_n256m_1gpu.err; echo "bench 1gpu rc=$?"
python -c "
import json; d=json.load(open('$out/bench_n256m_1gpu.json')); print('1gpu', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['invariants']['min_d2'], 'e2e %.4e'%d['e2e']['value'])"
fi
