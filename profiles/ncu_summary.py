"""Per-kernel summary of an .ncu-rep captured with --set full (averages over the profiled launches).
usage: python profiles/ncu_summary.py <report.ncu-rep> <out.json> [note]"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct", "launch__registers_per_thread": "registers",
    "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__waves_per_multiprocessor": "waves_per_sm",
    "launch__occupancy_limit_registers": "blocks_per_sm_limit_registers", "launch__occupancy_limit_shared_mem": "blocks_per_sm_limit_smem",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
}
acc = defaultdict(lambda: defaultdict(list))
name_i = h.index("Kernel Name")
for r in rows[2:]:
    k = r[name_i].split("(")[0].replace("void ", "").replace("unnamed>::", "").strip()
    for m, short in want.items():
        if m in h:
            i = h.index(m)
            try:
                acc[k][f"{short} [{units[i]}]"].append(float(r[i].replace(",", "")))
            except ValueError:
                pass
summary = {"source": note, "kernels": {k: dict({m: round(sum(v) / len(v), 3) for m, v in d.items()}, launches_profiled=len(next(iter(d.values()))))
                                       for k, d in acc.items()}}
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps(summary, indent=1)[:3000])
