#!/bin/bash
# round 2, GPU call 15 (1 GPU): reproduce the city blow-up of call 14 (pool exhaustion + replay in fleet mode) with tiny pools
mkdir -p gpurun_out/r2c15
BNX_DEBUG=1 BNX_INIT_LEAF_MB=8 BNX_GROW_MB=1 BNX_PEER_TIMEOUT_MS=120000 timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29621 \
  tools/city_fleet.py --steps 80 --check 20 --oracle-steps 4 --out gpurun_out/r2c15/city_3on1_tiny.json > /dev/null 2> gpurun_out/r2c15/city_3on1_tiny.err
echo "rc=$?" >> gpurun_out/r2c15/city_3on1_tiny.err
timeout 200 python tools/city_fleet.py --vehicles 3 --steps 80 --check 20 --oracle-steps 4 --out gpurun_out/r2c15/city_1gpu_3veh.json > /dev/null 2> gpurun_out/r2c15/city_1gpu_3veh.err
echo "rc=$?" >> gpurun_out/r2c15/city_1gpu_3veh.err
python tools/compare_city.py gpurun_out/r2c15/city_3on1_tiny.json gpurun_out/r2c15/city_1gpu_3veh.json > gpurun_out/r2c15/compare.json 2>&1
cat gpurun_out/r2c15/compare.json
