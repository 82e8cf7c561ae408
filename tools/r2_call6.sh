#!/bin/bash
# round 2, GPU call 6 (4 GPUs): multi-process sharded parity on real GPUs + the sharded bench at the driver's step count
mkdir -p gpurun_out/r2c6
nvidia-smi -L > gpurun_out/r2c6/gpus.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q --durations=8 > gpurun_out/r2c6/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c6/pytest.log
run() { # n steps warmup tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $1 --steps $2 --warmup $3 \
    > gpurun_out/r2c6/bench_n$1_$4.json 2> gpurun_out/r2c6/bench_n$1_$4.err
}
run 4 20 5 short
run 2 20 5 short
run 4 200 10 long
tail -3 gpurun_out/r2c6/pytest.log
