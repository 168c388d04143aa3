#!/bin/sh
# round-2 GPU call e (2 GPUs): slab runs with the crowded-cell flags in the ring: bit identity, 256M hash, timing
out=gpurun_out/r2e; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -k "two_rank or assign" > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/pytest.log
for mode in "262144 7" "4194304 9" "1048576 12 crowded"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 scripts/slab_worker.py $mode 2>&1 | grep SLAB | tee -a $out/slab_worker.log
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n256m_2gpu.json 2> $out/bench_n256m_2gpu.err; echo "bench 2gpu rc=$?"
python -c "
import json; d=json.load(open('$out/bench_n256m_2gpu.json')); print('2gpu', '%.4e'%d['value'], d['roofline']['ms_per_sweep'], d['invariants']['state_hash'], d['invariants']['min_d2'], 'e2e %.4e'%d['e2e']['value'])"
