#!/bin/bash
# round 2, GPU call 12 (1 GPU): apply pass through the bulk-copy engine (cp.async.bulk + mbarrier): parity, A/B bench, ncu
mkdir -p gpurun_out/r2c12
timeout 600 python -m pytest tests/test_gpu_map.py tests/test_gpu_golden.py "tests/test_gpu_fullsize.py::test_config3_lidar_200_scans" -x -q -k "not dense" --durations=3 > gpurun_out/r2c12/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c12/pytest.log
timeout 300 python bench.py --steps 300 --warmup 10 --no-dropin --no-cpu > gpurun_out/r2c12/bench_tma.json 2> gpurun_out/r2c12/bench_tma.err
BNX_APPLY_TMA=0 timeout 300 python bench.py --steps 300 --warmup 10 --no-dropin --no-cpu > gpurun_out/r2c12/bench_reg.json 2> gpurun_out/r2c12/bench_reg.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-dropin --no-cpu > gpurun_out/r2c12/bench_tma20.json 2> gpurun_out/r2c12/bench_tma20.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_apply_leaves" -s 6 -c 4 -o gpurun_out/r2c12/prof_apply python bench.py --steps 14 --warmup 3 --no-cpu --no-dropin > /dev/null 2> gpurun_out/r2c12/ncu.err
tail -3 gpurun_out/r2c12/pytest.log
