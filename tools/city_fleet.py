#!/usr/bin/env python
"""BASELINE.json config #5: a city-scale map (~1e9 active voxels) built by a FLEET of vehicles into ONE map sharded by root
key over the GPUs (one vehicle per rank, fleet steps: bnx_map_shard_set_fleet + bnx_map_shard_insert).

    python -m torch.distributed.run --nproc-per-node N tools/city_fleet.py --steps S --out profiles/r2_config5_city_nN.json
    python tools/city_fleet.py --vehicles V --steps S --out ...     # ONE GPU: the same V-vehicle workload, the vehicles'
                                                                    # scans inserted one after the other (reference order)

Vehicle v drives the serpentine city path of bonxai_b200.synth (2 m per step, streets 60 m apart) in its own district,
3 km to the side of vehicle v - 1. Every `--check` steps the order-independent digest of the whole map is recorded: the
N-GPU run, the 1-GPU run and (for the first steps) the CPU oracle must agree on it."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bonxai_b200 import synth  # noqa: E402

RES, MAX_RANGE, DISTRICT = 0.1, 50.0, 3000.0


def vehicle_scan(job):
    step, v = job
    pts, origin = synth.lidar_scan(step, speed=2.0, path="city", seed=7 + v)
    shift = np.float32([0.0, DISTRICT * v, 0.0])
    pts = pts.copy()
    pts[:, :3] += shift
    return pts, origin + shift


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1200)
    ap.add_argument("--vehicles", type=int, default=0, help="single process only: how many vehicles to interleave")
    ap.add_argument("--check", type=int, default=100)
    ap.add_argument("--oracle-steps", type=int, default=8)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    vehicles = world if world > 1 else max(1, args.vehicles)
    mine = [rank] if world > 1 else list(range(vehicles))
    import multiprocessing as mp
    procs = max(1, min(24, ((os.cpu_count() or 2) - 1) // world))
    pool = mp.get_context("fork").Pool(procs)  # fork before CUDA
    batch = args.check

    n_oracle = min(args.oracle_steps, args.steps)

    def gen(first):
        # the first batch ends where the oracle's steps end, so that checkpoint 0 can be compared with the CPU reference
        last = n_oracle if first == 0 and n_oracle else min(first + batch, args.steps)
        jobs = [(s, v) for s in range(first, last) for v in mine]
        return pool.map_async(vehicle_scan, jobs, chunksize=max(1, len(jobs) // (procs * 4)))

    # all scans are generated BEFORE anything is timed: generator processes that run next to the timed windows take the
    # cores the enqueueing threads need (one rank that is late to enqueue stalls every rank at the next exchange)
    batches, first = [], 0
    while first < args.steps:
        batches.append(gen(first).get())
        first += len(batches[-1]) // len(mine)
    oracle_jobs = [(s, v) for s in range(min(args.oracle_steps, args.steps)) for v in range(vehicles)] if rank == 0 else []
    oracle_scans = pool.map(vehicle_scan, oracle_jobs) if oracle_jobs else []

    pool.terminate()
    import torch
    from bonxai_b200 import capi
    torch.cuda.set_device(local % torch.cuda.device_count())
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    if world > 1:
        import torch.distributed as dist
        from bonxai_b200.sharded import ShardedMap
        same_gpu = torch.cuda.device_count() < world
        dist.init_process_group("gloo" if same_gpu else "nccl", **({} if same_gpu else {"device_id": torch.device("cuda", local)}))
        sm = ShardedMap(RES, bootstrap="host" if same_gpu else "nccl")
        gm = sm.map
    else:
        gm = capi.ProbabilisticMap(RES)
        gm.set_stream(stream.cuda_stream)

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cpu" if same_gpu else "cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(xs):
        if world == 1:
            return list(xs)
        t = torch.tensor(list(xs), dtype=torch.float64, device="cpu" if same_gpu else "cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    checkpoints, gpu_ms, done = [], 0.0, 0
    oracle_digest = None
    t_wall = time.perf_counter()
    while done < args.steps:
        scans = batches.pop(0)
        nsteps = len(scans) // len(mine)
        dev = [torch.from_numpy(p).cuda() for p, _ in scans]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(nsteps):
            if world > 1:
                p, o = scans[k]
                origins = np.array([o + np.float32([0.0, DISTRICT * (v - rank), 0.0]) for v in range(world)], np.float64)
                sm.insert_fleet(capi.DevPtr(dev[k].data_ptr()), len(p), 16, len(p), origins, MAX_RANGE, use_async=True)
            else:
                for j in range(len(mine)):
                    p, o = scans[k * len(mine) + j]
                    gm.insert_async(capi.DevPtr(dev[k * len(mine) + j].data_ptr()), o, MAX_RANGE, n=len(p), stride_bytes=16)
        e1.record(stream)
        gm.sync()
        ms = allmax(e0.elapsed_time(e1))
        gpu_ms += ms
        done += nsteps
        dig = sm.digest() if world > 1 else gm.digest()
        st = gm.grid().stats()
        live_b, mapped_b = allsum([st["leaves"] * 2304.0, float(st["mapped_bytes"])])
        cp = {"steps": done, "scans": done * vehicles, "active_cells": dig[2], "digest": [hex(dig[0]), hex(dig[1])], "ms_per_step": ms / nsteps,
              "leaf_bytes_live_GB": live_b / 2**30, "mapped_GB": mapped_b / 2**30}
        checkpoints.append(cp)
        if rank == 0:
            print(json.dumps(cp), file=sys.stderr, flush=True)
        if rank == 0 and oracle_digest is None and oracle_scans:
            import oracle
            kind = "reference" if oracle.available("reference") else "port"
            om = oracle.load(kind).map(RES)
            single = capi.ProbabilisticMap(RES)  # and the unsharded GPU map, scan after scan, for the same steps
            for p, o in oracle_scans:
                om.insert(p, o, MAX_RANGE)
                single.insert(p, o, MAX_RANGE)
            od, sd = om.digest(), single.digest()
            oracle_digest = {"steps": len(oracle_scans) // vehicles, "oracle": kind, "digest_oracle": list(map(hex, od[:2])),
                             "digest_single_gpu": list(map(hex, sd[:2])), "active_cells": od[2],
                             "this_run_equals_oracle": checkpoints[0]["steps"] == len(oracle_scans) // vehicles and
                                                       checkpoints[0]["digest"] == list(map(hex, od[:2])) and checkpoints[0]["active_cells"] == od[2],
                             "single_gpu_equals_oracle": sd == od}
            single.close()
            del om
    tt = gm.totals()
    N_all, U_all, V_all = allsum([tt["N"], tt["U"], tt["V"] + (tt["N"] if world > 1 else 0)])
    stats = sm.stats() if world > 1 else None
    if rank == 0:
        out = {"config": 5, "workload": f"city-scale fleet: {vehicles} vehicles x {args.steps} steps (131,072 pts/scan, 0.1 m, 50 m, 2 m/step, serpentine "
                                       f"streets, districts {DISTRICT:.0f} m apart) into ONE map" + (f" sharded by root key over {world} GPUs" if world > 1 else " on 1 GPU"),
               "n_gpus": world, "vehicles": vehicles, "steps": done, "scans": done * vehicles, "active_cells": checkpoints[-1]["active_cells"],
               "gpu_ms_per_step": gpu_ms / done, "points_per_s": N_all / gpu_ms * 1e3, "voxel_updates_per_s": U_all / gpu_ms * 1e3,
               "ray_visits_per_s": V_all / gpu_ms * 1e3, "mapped_over_live": checkpoints[-1]["mapped_GB"] / max(checkpoints[-1]["leaf_bytes_live_GB"], 1e-9),
               "slowest_window_vs_median": max(c["ms_per_step"] for c in checkpoints[1:] or checkpoints) /
                                           float(np.median([c["ms_per_step"] for c in checkpoints[1:] or checkpoints])),
               "first_steps_parity": oracle_digest, "pipeline": stats, "wall_s": time.perf_counter() - t_wall, "checkpoints": checkpoints}
        # the first checkpoint that coincides with the oracle's step count is compared here; the rest by diffing the
        # checkpoint digests of runs with different GPU counts (tools/compare_city.py)
        txt = json.dumps(out)
        print(txt, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(txt + "\n")
    if world > 1:
        dist.barrier()
        sm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
