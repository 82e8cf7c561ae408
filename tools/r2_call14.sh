#!/bin/bash
# round 2, GPU call 14 (8 GPUs): fleet bench at the driver's step count + the city-scale fleet run (8 vehicles, ~1e9 voxels)
mkdir -p gpurun_out/r2c14
BNX_BENCH_WATCHDOG=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c14/bench_n8_fleet.json 2> gpurun_out/r2c14/bench_n8_fleet.err
echo "rc=$?" >> gpurun_out/r2c14/bench_n8_fleet.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 \
  tools/city_fleet.py --steps 1200 --check 100 --oracle-steps 2 --out gpurun_out/r2c14/city_n8.json > /dev/null 2> gpurun_out/r2c14/city_n8.err
echo "rc=$?" >> gpurun_out/r2c14/city_n8.err
BNX_BENCH_WATCHDOG=100 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 \
  bench.py --gpus 8 --steps 20 --warmup 5 --workload dense-scan > gpurun_out/r2c14/bench_n8_dense.json 2> gpurun_out/r2c14/bench_n8_dense.err
echo "rc=$?" >> gpurun_out/r2c14/bench_n8_dense.err
nproc > gpurun_out/r2c14/nproc.txt
