#!/usr/bin/env python
"""Compares the checkpoint digests of two tools/city_fleet.py runs of the SAME fleet workload (e.g. N GPUs vs 1 GPU):
usage: python tools/compare_city.py a.json b.json"""
import json
import sys

a, b = (json.load(open(f)) for f in sys.argv[1:3])
assert a["vehicles"] == b["vehicles"], "different workloads"
ca = {c["steps"]: (c["digest"], c["active_cells"]) for c in a["checkpoints"]}
cb = {c["steps"]: (c["digest"], c["active_cells"]) for c in b["checkpoints"]}
common = sorted(set(ca) & set(cb))
bad = [s for s in common if ca[s] != cb[s]]
print(json.dumps({"a": sys.argv[1], "b": sys.argv[2], "vehicles": a["vehicles"], "gpus": [a["n_gpus"], b["n_gpus"]], "checkpoints_compared": len(common),
                  "last_step_compared": common[-1] if common else None, "active_cells_at_last": ca[common[-1]][1] if common else None,
                  "mismatching_steps": bad, "equal": not bad and bool(common)}))
sys.exit(1 if bad or not common else 0)
