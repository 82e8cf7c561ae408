#!/bin/bash
# round 2, GPU call 10 (2 GPUs): fleet bench on real GPUs after the overlap change, driver step count and a long run
mkdir -p gpurun_out/r2c10
BNX_BENCH_WATCHDOG=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29584 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c10/bench_n2_fleet.json 2> gpurun_out/r2c10/bench_n2_fleet.err
echo "rc=$?" >> gpurun_out/r2c10/bench_n2_fleet.err
BNX_BENCH_WATCHDOG=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29585 \
  bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2c10/bench_n2_fleet_long.json 2> gpurun_out/r2c10/bench_n2_fleet_long.err
echo "rc=$?" >> gpurun_out/r2c10/bench_n2_fleet_long.err
BNX_BENCH_WATCHDOG=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29586 \
  bench.py --gpus 2 --steps 20 --warmup 5 --workload dense-scan > gpurun_out/r2c10/bench_n2_dense.json 2> gpurun_out/r2c10/bench_n2_dense.err
echo "rc=$?" >> gpurun_out/r2c10/bench_n2_dense.err
