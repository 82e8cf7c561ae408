#!/bin/bash
# round 2, GPU call 9 (1 GPU): sharded suite after the overlap change; fleet bench with 4 ranks sharing the GPU, short and long
mkdir -p gpurun_out/r2c9
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --durations=5 > gpurun_out/r2c9/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c9/pytest.log
BNX_BENCH_WATCHDOG=280 BNX_PEER_TIMEOUT_MS=120000 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 \
  bench.py --gpus 4 --steps 150 --warmup 10 > gpurun_out/r2c9/bench_4on1_fleet_long.json 2> gpurun_out/r2c9/bench_4on1_fleet_long.err
echo "rc=$?" >> gpurun_out/r2c9/bench_4on1_fleet_long.err
BNX_BENCH_WATCHDOG=200 BNX_PEER_TIMEOUT_MS=120000 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29592 \
  tools/city_fleet.py --steps 60 --check 20 --oracle-steps 4 --out gpurun_out/r2c9/city_3on1.json > /dev/null 2> gpurun_out/r2c9/city_3on1.err
echo "rc=$?" >> gpurun_out/r2c9/city_3on1.err
timeout 240 python tools/city_fleet.py --vehicles 3 --steps 60 --check 20 --oracle-steps 4 --out gpurun_out/r2c9/city_1gpu_3veh.json > /dev/null 2> gpurun_out/r2c9/city_1gpu_3veh.err
echo "rc=$?" >> gpurun_out/r2c9/city_1gpu_3veh.err
tail -3 gpurun_out/r2c9/pytest.log
