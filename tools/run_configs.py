#!/usr/bin/env python
"""Runs the BASELINE.json configurations that are not the bench.py line and prints one JSON
object per configuration. Parity is checked where the CPU oracle finishes in seconds; larger sizes are
checked through size-independent properties (counts, round trips, digests across code paths).

    python tools/run_configs.py sweep [--max-log2 27]     config #2  VoxelGrid<float> create/update/value/forEachCell
    python tools/run_configs.py depth [--scans 20]        config #4  1280x800 depth camera, 0.01 m, 5 m
    python tools/run_configs.py city  [--cells 1e9]       config #5  city-scale LiDAR map on one GPU
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from bonxai_b200 import synth  # noqa: E402


def _events(torch):
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


# ------------------------------------------------------------------------------------------------ config #2
def sweep(args):
    import torch
    from bonxai_b200 import capi
    import oracle
    from workloads import digest

    port = oracle.load("port")
    stream = torch.cuda.Stream()  # a real stream: the NULL stream serialises the programmatic dependent launches
    for log2n in range(20, args.max_log2 + 1, 2 if args.max_log2 > 24 else 4):
        n = 1 << log2n
        for pattern in ("coherent_x", "coherent_z", "random"):
            # coordinates generated on the device for the big sizes (same formulas as bonxai_b200.synth)
            i = torch.arange(n, dtype=torch.int64, device="cuda")
            if pattern.startswith("coherent"):
                side = int(np.ceil(round(n ** (1.0 / 3.0), 9)))
                a, b, c = i % side, (i // side) % side, i // (side * side)
                xyz = torch.stack([a, b, c] if pattern.endswith("x") else [c, b, a], dim=1) - side // 2
            else:
                side = int(np.ceil((2.0 * n) ** (1.0 / 3.0)))
                g = torch.Generator(device="cuda").manual_seed(42)
                xyz = torch.randint(0, side, (n, 3), generator=g, device="cuda", dtype=torch.int64) - side // 2
            xyz = xyz.to(torch.int32).contiguous()
            vals = (i & 0xFFFF).to(torch.float32)
            was = torch.empty(n, dtype=torch.uint8, device="cuda")
            out = torch.zeros(n, dtype=torch.float32, device="cuda")
            found = torch.empty(n, dtype=torch.uint8, device="cuda")
            del i
            grid = capi.VoxelGrid(0.05, dtype=np.float32)
            grid.set_stream(stream.cuda_stream)
            res = {"config": 2, "n": n, "pattern": pattern}
            for op in ("create", "update", "value", "forEachCell"):
                e0, e1 = _events(torch)
                torch.cuda.synchronize()
                e0.record(stream)
                if op in ("create", "update"):
                    grid.set_values(capi.DevPtr(xyz.data_ptr()), capi.DevPtr(vals.data_ptr()), n=n, was_on=capi.DevPtr(was.data_ptr()))
                elif op == "value":
                    grid.get_values(capi.DevPtr(xyz.data_ptr()), n=n, values=capi.DevPtr(out.data_ptr()), found=capi.DevPtr(found.data_ptr()))
                else:
                    cnt = grid.active_count()
                    dx = torch.empty((cnt, 3), dtype=torch.int32, device="cuda")
                    dv = torch.empty(cnt, dtype=torch.float32, device="cuda")
                    torch.cuda.synchronize()
                    e0.record(stream)
                    got = grid.dump_device(capi.DevPtr(dx.data_ptr()), capi.DevPtr(dv.data_ptr()), cnt)
                    assert got == cnt
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                units = n if op != "forEachCell" else cnt
                res[op] = {"ms": ms, "Mops_per_s": units / ms / 1e3, "algorithmic_GBps": 20.0 * units / ms / 1e6}
                if op == "create":
                    uniq = grid.active_count()
                    res["active_cells"] = uniq
                    assert int(was.sum().item()) == n - uniq, "was_on must be false exactly once per distinct cell"
                elif op == "update":
                    assert bool(was.all().item())
                elif op == "value":
                    assert bool(found.all().item())
                    if pattern != "random":
                        assert torch.equal(out, vals)
            if log2n <= 20:  # CPU oracle parity by dump digest
                og = port.grid(0.05)
                hx, hv = xyz.cpu().numpy(), vals.cpu().numpy()
                t0 = time.perf_counter()
                og.set_values(hx, hv)
                res["cpu_create_ms"] = 1e3 * (time.perf_counter() - t0)
                t0 = time.perf_counter()
                og.set_values(hx, hv)
                res["cpu_update_ms"] = 1e3 * (time.perf_counter() - t0)
                ox, ov = og.dump(sort=False)
                assert digest(ox, ov) == digest(dx.cpu().numpy(), dv.cpu().numpy().view(np.uint32)), "dump differs from the oracle"
                res["parity"] = "dump digest == oracle"
            else:
                res["parity"] = "properties (was_on counts, value round trip, dump count)"
            res["stats"] = grid.stats()
            print(json.dumps(res), flush=True)
            del grid, xyz, vals, was, out, found, dx, dv
            torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ config #4
def depth(args):
    import torch
    from bonxai_b200 import capi
    import oracle
    from workloads import digest

    port = oracle.load("port")
    scans = [synth.depth_scan(s) for s in range(args.scans)]
    dev = [torch.from_numpy(p).cuda() for p, _ in scans]
    stream = torch.cuda.Stream()  # a real stream: the NULL stream serialises the programmatic dependent launches
    # parity: the first scans against the CPU oracle (about 5 s of CPU each)
    gm, om = capi.ProbabilisticMap(0.01), port.map(0.01)
    cpu_s, tot = [], dict(N=0, E=0, V=0, U=0)
    for k in range(min(args.parity_scans, args.scans)):
        gm.insert(scans[k][0], scans[k][1], 5.0)
        om.insert(scans[k][0], scans[k][1], 5.0)
        cpu_s.append(om.last_insert_seconds())
        assert digest(*gm.dump(sort=False)) == digest(*om.dump(sort=False)), f"depth scan {k}: GPU map differs from the oracle"
        oc, gc = om.counters(), gm.counters()
        assert (oc["N"], oc["E"], oc["V"], oc["U"]) == (gc["N"], gc["E"], gc["V"], gc["U"])
    del gm, om
    m = capi.ProbabilisticMap(0.01)
    m.set_stream(stream.cuda_stream)
    for k in range(2):
        m.insert_async(capi.DevPtr(dev[k].data_ptr()), scans[k][1], 5.0, n=len(scans[k][0]), stride_bytes=16)
    m.sync()
    t0 = m.totals()
    e0, e1 = _events(torch)
    e0.record(stream)
    for k in range(2, args.scans):
        m.insert_async(capi.DevPtr(dev[k].data_ptr()), scans[k][1], 5.0, n=len(scans[k][0]), stride_bytes=16)
    e1.record(stream)
    m.sync()
    ms = e0.elapsed_time(e1)
    t1 = m.totals()
    n_timed = args.scans - 2
    d = {k: t1[k] - t0[k] for k in t1}
    print(json.dumps({"config": 4, "workload": "depth 1280x800 (1,024,000 pts/scan), res 0.01 m, max_range 5 m", "scans_timed": n_timed,
                      "ms_per_scan": ms / n_timed, "points_per_s": d["N"] / ms * 1e3, "voxel_updates_per_s": d["U"] / ms * 1e3,
                      "ray_visits_per_s": d["V"] / ms * 1e3, "visits_per_scan": d["V"] / n_timed, "updates_per_scan": d["U"] / n_timed,
                      "active_cells": m.active_count(), "parity": f"dump digest + counters == oracle on scans 0..{len(cpu_s) - 1}",
                      "cpu_reference_port_ms_per_scan": 1e3 * float(np.mean(cpu_s)) if cpu_s else None, "stats": m.grid().stats()}), flush=True)


# ------------------------------------------------------------------------------------------------ config #5
def _city_scan(s):
    return synth.lidar_scan(s, speed=2.0, path="city")


def city(args):
    import multiprocessing as mp
    target = float(args.cells)
    batch = 512
    pool = mp.get_context("fork").Pool(max(1, min(32, (os.cpu_count() or 2) - 1)))
    first = pool.map(_city_scan, range(batch), chunksize=8)  # fork before CUDA

    import torch
    from bonxai_b200 import capi
    import oracle
    from workloads import digest

    stream = torch.cuda.Stream()  # a real stream: the NULL stream serialises the programmatic dependent launches
    m = capi.ProbabilisticMap(0.1)
    m.set_stream(stream.cuda_stream)
    om = oracle.load("port").map(0.1)
    gpu_ms, scans_done, checked = 0.0, 0, 0
    t_wall = time.perf_counter()
    cur = first
    log = []
    while True:
        nxt = pool.map_async(_city_scan, range(scans_done + len(cur), scans_done + len(cur) + batch), chunksize=8)
        dev = [torch.from_numpy(p).cuda() for p, _ in cur]
        e0, e1 = _events(torch)
        e0.record(stream)
        for (p, o), d in zip(cur, dev):
            m.insert_async(capi.DevPtr(d.data_ptr()), o, 50.0, n=len(p), stride_bytes=16)
        e1.record(stream)
        m.sync()
        gpu_ms += e0.elapsed_time(e1)
        if scans_done == 0:  # oracle parity on the first 64 scans of the city path
            m2 = capi.ProbabilisticMap(0.1)
            for p, o in cur[:64]:
                m2.insert(p, o, 50.0)
                om.insert(p, o, 50.0)
            assert digest(*m2.dump(sort=False)) == digest(*om.dump(sort=False)), "city path: GPU map differs from the oracle after 64 scans"
            checked = 64
            del m2
        scans_done += len(cur)
        active = m.active_count()
        st = m.grid().stats()
        log.append({"scans": scans_done, "active_cells": active, "leaves": st["leaves"], "roots": st["roots"], "mapped_GB": st["mapped_bytes"] / 2**30,
                    "gpu_ms_per_scan_cum": gpu_ms / scans_done})
        print(json.dumps(log[-1]), file=sys.stderr, flush=True)
        if active >= target or scans_done >= args.max_scans:
            break
        cur = nxt.get()
    pool.terminate()
    tt = m.totals()
    # the pipelined map must equal a map built with synchronous calls: compare digests on a re-run prefix? Too long;
    # instead: occupied-voxel list size and a full-dump digest are reported for cross-run / cross-GPU-count comparison
    xyz, w = m.dump(sort=False)
    print(json.dumps({"config": 5, "workload": "city-scale LiDAR (131,072 pts/scan, 0.1 m, 50 m, 2 m/scan, serpentine streets)",
                      "scans": scans_done, "active_cells": int(len(xyz)), "digest": digest(xyz, w), "gpu_ms_per_scan": gpu_ms / scans_done,
                      "points_per_s": tt["N"] / gpu_ms * 1e3, "voxel_updates_per_s": tt["U"] / gpu_ms * 1e3, "ray_visits_per_s": tt["V"] / gpu_ms * 1e3,
                      "parity": f"dump digest == oracle after the first {checked} scans", "wall_s": time.perf_counter() - t_wall,
                      "stats": m.grid().stats(), "growth": log[:: max(1, len(log) // 8)]}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("sweep")
    s.add_argument("--max-log2", type=int, default=24)
    d = sub.add_parser("depth")
    d.add_argument("--scans", type=int, default=20)
    d.add_argument("--parity-scans", type=int, default=2)
    c = sub.add_parser("city")
    c.add_argument("--cells", default="1e9")
    c.add_argument("--max-scans", type=int, default=8192)
    args = ap.parse_args()
    {"sweep": sweep, "depth": depth, "city": city}[args.cmd](args)


if __name__ == "__main__":
    main()
