#!/bin/bash
# round 2, GPU call 19 (1 GPU, last minutes of the budget): pool growth behind a running pipeline with a decaying growth estimate
mkdir -p gpurun_out/r2c19
BNX_INIT_LEAF_MB=256 timeout 100 python tools/city_fleet.py --vehicles 2 --steps 500 --check 50 --oracle-steps 0 --out gpurun_out/r2c19/city_1gpu_2veh_growth.json > /dev/null 2> gpurun_out/r2c19/city.err
echo "rc=$?" >> gpurun_out/r2c19/city.err
grep "^{" gpurun_out/r2c19/city.err | cut -c1-60,150-260
