#!/bin/sh
# 8 GPUs: does the number of NCCL p2p channels (one CTA each, next to the sweep kernels) matter?  device-timed value only
N=$1; out=gpurun_out/r4l_$N; mkdir -p $out
for ch in default 2 1; do
  if [ "$ch" = "default" ]; then unset NCCL_MAX_P2P_NCHANNELS; else export NCCL_MAX_P2P_NCHANNELS=$ch; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29676 bench.py --gpus $N --steps 4 --warmup 3 --no-e2e > $out/bench_$ch.json 2> $out/bench_$ch.err; echo "rc=$?"
  python -c "
import json; d=json.load(open('$out/bench_$ch.json')); print('ch=$ch', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['clocks']['sm_mhz'])"
done
