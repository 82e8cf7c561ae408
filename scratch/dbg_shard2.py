import sys, numpy as np
sys.path.insert(0, "/root/repo")
from bonxai_b200 import capi, synth
from bonxai_b200.sharded import LocalShardGroup
import oracle
port = oracle.load("port")
for world in (2,):
    g, om = LocalShardGroup(0.1, world), port.map(0.1)
    pts = np.array([[30.03, 0.05, 0.05],[0.05, 25.0, 0.05]], np.float32)
    g.insert(pts, [0.05,0.05,0.05], 100.0); om.insert(pts, [0.05,0.05,0.05], 100.0)
    gx, gw = g.dump(); ox, ow = om.dump()
    gs = set(map(tuple, gx))
    missing = np.array([c for c in map(tuple, ox) if c not in gs])
    print("world", world, "cells", len(gx), len(ox), "missing:", missing.tolist()[:80])
    for r, s in enumerate(g.shards):
        x, w = s.map.dump()
        print("shard", r, "cells", len(x), "x-range by root:", sorted(set((x[:,0]>>5).tolist())), sorted(set((x[:,1]>>5).tolist())), s.map.counters())
        print(" send2 hdr counts", s.send2[:,0,0].tolist(), "recv2 hdr", s.recv2[:,0,0].tolist(), "send1", s.send1[:,0,0].tolist(), "recv1", s.recv1[:,0,0].tolist())
