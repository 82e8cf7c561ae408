#!/bin/bash
# round 2, GPU call 13 (4 GPUs): fleet bench at the driver's step count (both arms), dense-scan bench, city fleet (4 vehicles)
mkdir -p gpurun_out/r2c13
run() { # tag, args...
  tag=$1; shift
  BNX_BENCH_WATCHDOG=170 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29601 "$@" \
    > gpurun_out/r2c13/$tag.json 2> gpurun_out/r2c13/$tag.err
  echo "rc=$?" >> gpurun_out/r2c13/$tag.err
}
run bench_n4_fleet bench.py --gpus 4 --steps 20 --warmup 5
run bench_n4_dense bench.py --gpus 4 --steps 20 --warmup 5 --workload dense-scan
run bench_n4_ref bench.py --gpus 4 --steps 20 --warmup 5 --impl reference
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29602 \
  tools/city_fleet.py --steps 1200 --check 100 --oracle-steps 4 --out gpurun_out/r2c13/city_n4.json > /dev/null 2> gpurun_out/r2c13/city_n4.err
echo "rc=$?" >> gpurun_out/r2c13/city_n4.err
tail -2 gpurun_out/r2c13/city_n4.err
