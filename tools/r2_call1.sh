#!/bin/bash
# round 2, GPU call 1 (1 GPU): full -m gpu suite, N=1 bench at the driver's step count, the sharded bench with 8 ranks
# SHARING the GPU (protocol + counters only: the times mean nothing), VMM probe
mkdir -p gpurun_out/r2c1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2c1/smi.txt
./build/vmm_probe > gpurun_out/r2c1/vmm_probe.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2c1/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c1/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1/bench_n1.json 2> gpurun_out/r2c1/bench_n1.err
BNX_PEER_TIMEOUT_MS=120000 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c1/bench_8on1.json 2> gpurun_out/r2c1/bench_8on1.err
echo "8on1 rc=$?" >> gpurun_out/r2c1/bench_8on1.err
tail -3 gpurun_out/r2c1/pytest.log
