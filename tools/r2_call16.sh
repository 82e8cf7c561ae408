#!/bin/bash
# round 2, GPU call 16 (1 GPU): the 8-vehicle city run with 8 ranks SHARING the GPU vs the same workload on an unsharded map
mkdir -p gpurun_out/r2c16
BNX_DEBUG=1 BNX_PEER_TIMEOUT_MS=300000 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 \
  tools/city_fleet.py --steps 560 --check 40 --oracle-steps 0 --out gpurun_out/r2c16/city_8on1.json > /dev/null 2> gpurun_out/r2c16/city_8on1.err
echo "rc=$?" >> gpurun_out/r2c16/city_8on1.err
timeout 300 python tools/city_fleet.py --vehicles 8 --steps 560 --check 40 --oracle-steps 0 --out gpurun_out/r2c16/city_1gpu_8veh.json > /dev/null 2> gpurun_out/r2c16/city_1gpu_8veh.err
echo "rc=$?" >> gpurun_out/r2c16/city_1gpu_8veh.err
python tools/compare_city.py gpurun_out/r2c16/city_8on1.json gpurun_out/r2c16/city_1gpu_8veh.json > gpurun_out/r2c16/compare.json 2>&1
cat gpurun_out/r2c16/compare.json
