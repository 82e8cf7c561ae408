#!/bin/bash
# round 2, GPU call 17 (8 GPUs): the city-scale fleet run again (8 vehicles x 1200 steps) with window-sized head-room at the drains
mkdir -p gpurun_out/r2c17
BNX_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 \
  tools/city_fleet.py --steps 1200 --check 100 --oracle-steps 2 --out gpurun_out/r2c17/city_n8.json > /dev/null 2> gpurun_out/r2c17/city_n8.err
echo "rc=$?" >> gpurun_out/r2c17/city_n8.err
grep "^{" gpurun_out/r2c17/city_n8.err | tail -3 | cut -c1-200
