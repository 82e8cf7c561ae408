#!/bin/bash
# round 2, GPU call 20 (1 GPU): ncu --set full of the four scan kernels of the final build (profiles/r2_ncu_summary.json)
mkdir -p gpurun_out/r2c20
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"k_classify|k_resolve|k_mark|k_apply_leaves" -s 24 -c 8 -o gpurun_out/r2c20/prof python bench.py --steps 12 --warmup 3 --no-cpu --no-dropin > /dev/null 2> gpurun_out/r2c20/ncu.err
ls -la gpurun_out/r2c20
