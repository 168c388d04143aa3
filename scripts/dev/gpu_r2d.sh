#!/bin/sh
# round-2 GPU call d: tests, equation-of-state deliverable (config 3), 256M on one GPU through torchrun
out=gpurun_out/r2d; mkdir -p $out
python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest.log
python scripts/eos_run.py --n 16777216 --phi 0.716 --seeds 8 --burn 20000 --samples 20 --stride 100 --out $out/eos_n16m_phi0.716.json > $out/eos_n16m.log 2>&1; echo "eos 16M rc=$?"; tail -c 900 $out/eos_n16m.log
python scripts/eos_run.py --n 65536 --phi 0.716 --seeds 8 --burn 20000 --samples 20 --stride 100 --out $out/eos_n64k_phi0.716.json > $out/eos_n64k.log 2>&1; echo "eos 64k rc=$?"
python scripts/eos_run.py --n 4096 --phi 0.716 --seeds 8 --burn 20000 --samples 20 --stride 100 --oracle-n 4096 --out $out/eos_n4096_phi0.716.json > $out/eos_n4096.log 2>&1; echo "eos 4096 rc=$?"; tail -c 600 $out/eos_n4096.log
python scripts/eos_run.py --n 16777216 --phi 0.70 --seeds 4 --burn 20000 --samples 20 --stride 100 --out $out/eos_n16m_phi0.70.json > $out/eos_n16m_070.log 2>&1; echo "eos 16M 0.70 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 1 --steps 3 --warmup 3 > $out/bench_n256m_1gpu.json 2> $out/bench_n256m_1gpu.err; echo "bench 256M rc=$?"; head -c 1200 $out/bench_n256m_1gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29656 bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $out/bench_ref_torchrun.json 2> $out/bench_ref.err; echo "ref arm rc=$?"; head -c 600 $out/bench_ref_torchrun.json
