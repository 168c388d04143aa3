#!/bin/sh
# 256M on ONE GPU through torch.distributed.run (the driver's N = 1 line of the scaling series), kernel v16
out=gpurun_out/r4g; mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29673 bench.py --gpus 1 --steps 3 --warmup 3 > $out/bench_n256m_1gpu.json 2> $out/bench_n256m_1gpu.err; echo "bench 1gpu rc=$?"
tail -3 $out/bench_n256m_1gpu.err
python -c "
import json; d=json.load(open('$out/bench_n256m_1gpu.json')); print('1gpu', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['invariants']['min_d2'], d['e2e'])"
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
