#!/bin/bash
# round 2, GPU call 2 (1 GPU): dense marking window — parity (map tests in both markings, golden, config 3), bench with
# the dense and with the sparse marks, ncu of the scan kernels
mkdir -p gpurun_out/r2c2
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_golden.py tests/test_gpu_prestep.py tests/test_gpu_publish.py tests/test_gpu_dropin.py \
  "tests/test_gpu_fullsize.py::test_config3_lidar_200_scans" "tests/test_gpu_fullsize.py::test_config4_depth_full_size" -x -q --durations=10 > gpurun_out/r2c2/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c2/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c2/bench_dense.json 2> gpurun_out/r2c2/bench_dense.err
BNX_DENSE=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-dropin > gpurun_out/r2c2/bench_sparse.json 2> gpurun_out/r2c2/bench_sparse.err
timeout 600 python bench.py --steps 300 --warmup 10 --no-dropin --no-cpu > gpurun_out/r2c2/bench_dense300.json 2> gpurun_out/r2c2/bench_dense300.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_classify|k_resolve|k_mark|k_apply" -s 24 -c 8 -o gpurun_out/r2c2/prof python bench.py --steps 14 --warmup 3 --no-cpu --no-dropin > /dev/null 2> gpurun_out/r2c2/ncu.err
tail -3 gpurun_out/r2c2/pytest.log
