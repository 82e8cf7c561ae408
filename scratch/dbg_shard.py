import sys, numpy as np
sys.path.insert(0, "/root/repo")
from bonxai_b200 import capi, synth
from bonxai_b200.sharded import LocalShardGroup
import oracle
port = oracle.load("port")
world = 2
g, om = LocalShardGroup(0.1, world), port.map(0.1)
pts, origin = synth.lidar_scan(0, beams=32, azimuths=1024)
g.insert(pts, origin, 40.0); om.insert(pts, origin, 40.0)
gx, gw = g.dump(); ox, ow = om.dump()
print("attempts", g.attempts, "cells", len(gx), len(ox), [s.map.counters() for s in g.shards], om.counters())
gs = set(map(tuple, gx)); os_ = [tuple(c) for c in ox]
missing = np.array([c for c in os_ if c not in gs])
print("missing", len(missing))
if len(missing):
    r = np.linalg.norm((missing - np.floor(origin*10)), axis=1)
    print("missing dist from origin (cells): min/median/max", r.min(), np.median(r), r.max())
    print("missing z range", missing[:,2].min(), missing[:,2].max())
    # leaf-level: how many leaves entirely missing
    ml = set(map(tuple, missing >> 3)); gl = set(map(tuple, gx >> 3))
    print("missing leaves", len(ml), "of which absent from gpu map entirely", len(ml - gl))
    for s in g.shards: print(s.map.grid().stats())
