#!/bin/bash
# round 2, GPU call 5 (1 GPU): dense marking v4 (deferred test-before-set, bitmap -> list kernel, hinted resumable apply)
mkdir -p gpurun_out/r2c5
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_golden.py -x -q --durations=5 > gpurun_out/r2c5/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c5/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c5/bench_dense.json 2> gpurun_out/r2c5/bench_dense.err
timeout 600 python bench.py --steps 300 --warmup 10 --no-dropin --no-cpu > gpurun_out/r2c5/bench_dense300.json 2> gpurun_out/r2c5/bench_dense300.err
for v in seg16 seg64 mb8 mb4; do
  BNX_LIB=$PWD/build/variants/libbonxai_b200_$v.so timeout 300 python bench.py --steps 100 --warmup 10 --no-dropin --no-cpu > gpurun_out/r2c5/bench_$v.json 2> gpurun_out/r2c5/bench_$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_resolve|k_mark|k_apply|k_list" -s 24 -c 8 -o gpurun_out/r2c5/prof python bench.py --steps 14 --warmup 3 --no-cpu --no-dropin > /dev/null 2> gpurun_out/r2c5/ncu.err
tail -3 gpurun_out/r2c5/pytest.log
