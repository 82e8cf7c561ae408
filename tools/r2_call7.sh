#!/bin/bash
# round 2, GPU call 7 (1 GPU): fleet step parity (LocalShardGroup), long sharded runs with 4 ranks sharing the GPU (hang hunt)
mkdir -p gpurun_out/r2c7
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -k "fleet or infinite or equals_oracle" --durations=5 > gpurun_out/r2c7/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c7/pytest.log
BNX_BENCH_WATCHDOG=240 BNX_PEER_TIMEOUT_MS=120000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 4 --steps 150 --warmup 10 --workload dense-scan > gpurun_out/r2c7/bench_4on1_long.json 2> gpurun_out/r2c7/bench_4on1_long.err
echo "rc=$?" >> gpurun_out/r2c7/bench_4on1_long.err
BNX_BENCH_WATCHDOG=240 BNX_PEER_TIMEOUT_MS=120000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29582 \
  bench.py --gpus 4 --steps 20 --warmup 5 --workload fleet > gpurun_out/r2c7/bench_4on1_fleet.json 2> gpurun_out/r2c7/bench_4on1_fleet.err
echo "rc=$?" >> gpurun_out/r2c7/bench_4on1_fleet.err
tail -3 gpurun_out/r2c7/pytest.log
