#!/bin/bash
# round 2, GPU call 8 (2 GPUs): long pipelined sharded run on real GPUs (the 4-GPU 200-step run of call 6 hung), with a watchdog
mkdir -p gpurun_out/r2c8
BNX_DEBUG=1 BNX_BENCH_WATCHDOG=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 \
  bench.py --gpus 2 --steps 200 --warmup 10 --workload dense-scan > gpurun_out/r2c8/bench_n2_long.json 2> gpurun_out/r2c8/bench_n2_long.err
echo "rc=$?" >> gpurun_out/r2c8/bench_n2_long.err
BNX_BENCH_WATCHDOG=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29584 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c8/bench_n2_fleet.json 2> gpurun_out/r2c8/bench_n2_fleet.err
echo "rc=$?" >> gpurun_out/r2c8/bench_n2_fleet.err
tail -2 gpurun_out/r2c8/bench_n2_long.err
