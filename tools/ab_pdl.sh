# LiDAR bench with programmatic dependent launch off / on (MODES="0 1"); optional first arg: pytest selection
if [ -n "$1" ]; then python -m pytest $1 -m gpu -x -q 2>&1 | tail -3; fi
for m in ${MODES:-1}; do BNX_PDL=$m python bench.py --steps 300 --warmup 10 --no-cpu > gpurun_out/bench_pdl_$m.json 2> gpurun_out/bench_pdl_$m.err; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_pdl_$m.json") if l.startswith("{")][-1])
print("pdl=$m", round(d["value"]/1e6), d["ms_per_step"]*1e3, d["phase_us_per_scan"], d["e2e"]["ms_per_step"]*1e3, d["sync_call"]["ms_per_step"]*1e3)
PY
done
