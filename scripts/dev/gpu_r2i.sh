#!/bin/sh
# round-2 GPU call i (8 GPUs): slab bit-identity at 8 ranks, N = 2^28 strong-scaling line with the state hash
out=gpurun_out/r2i; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
for mode in "4194304 9" "4194304 12 crowded"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 scripts/slab_worker.py $mode 2>&1 | grep SLAB | tee -a $out/slab_worker_8gpu.log
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 8 --steps 3 --warmup 3 > $out/bench_n256m_8gpu.json 2> $out/bench_n256m_8gpu.err; echo "bench 8gpu rc=$?"
python - <<'PY'
import json
t=open('gpurun_out/r2i/bench_n256m_8gpu.json').read()
d=json.loads(t[t.index('{"metric'):].splitlines()[0])
print('8gpu', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['invariants']['min_d2'], 'e2e %.4e'%d['e2e']['value'])
PY
