#!/bin/sh
# final kernel on N GPUs ($1), 256M, K = $2 timed steps (default 10): value, e2e of a longer job stream, state hash
N=$1; K=${2:-10}; out=gpurun_out/r4k_$N; mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29675 bench.py --gpus $N --steps $K --warmup 3 > $out/bench_n256m_${N}gpu.json 2> $out/bench_n256m_${N}gpu.err; echo "bench ${N}gpu rc=$?"
python -c "
import json; d=json.load(open('$out/bench_n256m_${N}gpu.json')); print('${N}gpu', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['invariants']['state_hash_after_sweeps'], 'e2e %.4e'%d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step'])"
