# round-end measurement set on one GPU: default bench (both arms), ncu launch list, ncu --set full of the four scan kernels
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 64 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 30 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_classify|k_resolve|k_mark|k_apply_leaves" -s 24 -c 8 -o gpurun_out/final_prof python bench.py --steps 14 --warmup 3 --no-cpu > /dev/null 2>&1
tail -c 400 gpurun_out/final_bench.err; ls -la gpurun_out | tail -6
