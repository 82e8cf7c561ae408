#!/bin/bash
# round 2, GPU call 11 (1 GPU): aggregated atomics — parity (map + sharded suites) and the 1-GPU bench
mkdir -p gpurun_out/r2c11
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_sharded.py tests/test_gpu_golden.py -x -q --durations=5 > gpurun_out/r2c11/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c11/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c11/bench_n1.json 2> gpurun_out/r2c11/bench_n1.err
timeout 600 python bench.py --steps 300 --warmup 10 --no-dropin --no-cpu > gpurun_out/r2c11/bench_n1_300.json 2> gpurun_out/r2c11/bench_n1_300.err
tail -3 gpurun_out/r2c11/pytest.log
