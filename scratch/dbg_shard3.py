import sys, numpy as np
sys.path.insert(0, "/root/repo")
from bonxai_b200 import capi, synth
from bonxai_b200.sharded import LocalShardGroup
import oracle
port = oracle.load("port")
world = 2
for nb, na in ((2, 64), (4, 256), (8, 512), (32, 1024)):
    g, om = LocalShardGroup(0.1, world, cap_leaves=1<<16), port.map(0.1)
    pts, origin = synth.lidar_scan(0, beams=nb, azimuths=na)
    g.insert(pts, origin, 40.0); om.insert(pts, origin, 40.0)
    gx, gw = g.dump(); ox, ow = om.dump()
    print(nb, na, "cells", len(gx), len(ox), "attempts", g.attempts)
    if len(gx) != len(ox):
        gs = set(map(tuple, gx))
        missing = np.array([c for c in map(tuple, ox) if c not in gs])
        owned = [set(map(tuple, (s.map.dump()[0] >> 5))) for s in g.shards]
        mroot = [tuple(c) for c in (missing >> 5)]
        for r in range(world):
            print("  missing cells in roots owned by", r, sum(1 for c in mroot if c in owned[r]))
        print("  in unknown roots", sum(1 for c in mroot if not any(c in o for o in owned)))
        for r, s in enumerate(g.shards):
            print("  shard", r, "send2", s.send2[:,0,0].tolist(), "recv2", s.recv2[:,0,0].tolist(), s.map.counters())
            # leaves in recv2 records vs leaves present
            for src in range(world):
                cnt = int(s.recv2[src,0,0])
                recs = s.recv2[src,1:cnt+1].cpu().numpy()
                leaves = set(map(tuple, (recs[:, :3] >> 3)))
                have = set(map(tuple, (s.map.dump()[0] >> 3)))
                masks = recs[:, 4:].view(np.uint64)
                nbits = int(sum(bin(int(v)).count("1") for v in masks.ravel())) if cnt else 0
                print("    from", src, "records", cnt, "distinct leaves", len(leaves), "present in map", len(leaves & have), "bits", nbits)
        break
