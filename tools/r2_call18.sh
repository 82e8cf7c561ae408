#!/bin/bash
# round 2, GPU call 18 (1 GPU): final validation — the whole -m gpu suite, the 1-GPU bench at the driver's step count, the
# unsharded reference of the 8-vehicle city run (digests vs the 8-GPU run), launch list
mkdir -p gpurun_out/r2c18
timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2c18/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c18/pytest.log
timeout 100 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c18/bench_n1.json 2> gpurun_out/r2c18/bench_n1.err
timeout 170 python tools/city_fleet.py --vehicles 8 --steps 1200 --check 100 --oracle-steps 2 --out gpurun_out/r2c18/city_1gpu_8veh.json > /dev/null 2> gpurun_out/r2c18/city_1gpu_8veh.err
echo "rc=$?" >> gpurun_out/r2c18/city_1gpu_8veh.err
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r2c18/launches.csv python bench.py --steps 12 --warmup 3 --no-cpu --no-dropin > /dev/null 2>&1
tail -3 gpurun_out/r2c18/pytest.log
