#!/usr/bin/env python
"""bench.py — insertPointCloud throughput on the synthetic 64-beam LiDAR sequence (BASELINE.json config #3).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A step is ONE scan (131,072 points, 0.1 m voxels, max_range 50 m, sensor moving 1 m per scan) inserted into
the running map. Printed: ONE JSON line (see DESIGN.md §6 for every field).

  value     points/s with the scans already resident in HBM (device-pointer C-ABI call per scan)
  e2e       points/s through the same C-ABI call with HOST (pinned) buffers: H2D copy of every scan and the
            D2H status/counter read-back are inside the timed region
  roofline  dominant kernel: algorithmic bytes per scan (16*N + 8*U) / its CUDA-event time
  cpu_baseline  the unmodified reference (oracle/_ref) on ONE host core, first scans of the same sequence

--impl reference times the reference's own CPU implementation (oracle/_ref, else the plain-C port) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bonxai_b200 import synth  # noqa: E402

RES, MAX_RANGE, BEAMS, AZ = 0.1, 50.0, 64, 2048
N_PTS = BEAMS * AZ
WORKLOAD = "lidar64x2048_seq(131072 pts/scan, res 0.1 m, max_range 50 m, 1 m/scan)"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic(phase: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the phase's kernel, from the committed ncu --set full
    capture of the current kernels (profiles/r2_ncu_summary.json, tools/r2_call20.sh); None if that file has no entry."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")) as f:
            k = json.load(f)["kernels"]
        name = {"classify": "k_classify<0, 1, 1>", "resolve": "k_resolve<0, 0>", "mark": "k_mark<0>", "apply": "k_apply_leaves<0>"}[phase]
        e = k[name]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}
        tot = 0.0
        for key, v in e.items():
            if key.startswith("dram_read") or key.startswith("dram_write"):
                tot += v * scale[key.split("[")[1].rstrip("]")]
        return tot
    except Exception:
        return None


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region. NVML (what nvidia-smi reads) polled every 2 ms from a
    thread, because the timed region lasts tens of milliseconds — shorter than one `nvidia-smi -lms` period."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.reasons, self.power = [], set(), []
        self.smax = None
        self.stop_flag = threading.Event()
        self.thread = None

    def resume(self):
        self.sm, self.power = [], []
        self.reasons = set()
        self.paused = False

    def start(self, paused=False):
        self.paused = paused
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical GPUs: honour CUDA_VISIBLE_DEVICES if it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return

        def pump():
            while not self.stop_flag.is_set():
                if self.paused:
                    time.sleep(0.002)
                    continue
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                time.sleep(0.002)

        try:  # the first queries of a process are slow (lazy initialisation inside NVML): do them before anything is timed
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            pynvml.nvmlDeviceGetPowerUsage(h)
        except Exception:
            pass
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag.set()
        self.thread.join(timeout=1)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smax, "samples": len(self.sm),
                "power_w_max": max(self.power) if self.power else None, "reasons": sorted(self.reasons)}


def gen_scans(first: int, count: int):
    """the scans of the sequence, generated on the host cores in parallel (fork: call before CUDA is initialised)"""
    import multiprocessing as mp
    procs = min(count, max(1, (os.cpu_count() or 2) - 1), 48)
    if procs <= 1 or count < 4:
        return [synth.lidar_scan(s) for s in range(first, first + count)]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(synth.lidar_scan, range(first, first + count), chunksize=max(1, count // (procs * 4)))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------
def load_cpu_oracle():
    import oracle
    if oracle.available("reference"):
        return oracle.load("reference"), "reference"
    return oracle.load("port"), "port"


def time_cpu(scans, warmup: int):
    """single-threaded reference insertPointCloud over `scans`; the clock brackets the call only
    (bonxai_map/benchmark/benchmark_kitti.cpp:146-152). Returns (seconds over timed scans, n timed)."""
    lib, kind = load_cpu_oracle()
    m = lib.map(RES)
    total = 0.0
    for i, (pts, origin) in enumerate(scans):
        m.insert(pts, origin, MAX_RANGE)
        if i >= warmup:
            total += m.last_insert_seconds()
    from bonxai_b200 import capi
    return total, len(scans) - warmup, kind, capi.digest_of_dump(*m.dump())


def run_dropin(scans, warmup: int):
    """e2e through the drop-in C++ headers: tools/cpp/dropin_bench.cpp (built against include/ + the library) inserts the
    same scans with the reference's own call, Bonxai::ProbabilisticMap::insertPointCloud(std::vector<PointXYZ>, origin,
    max_range), synchronously, from a pageable std::vector and from a PinnedAllocator vector, and with the publisher's
    post-step after every insert (bonxai_ros/src/bonxai_server.cpp:176-186,217-251). Separate process, own CUDA context."""
    import tempfile
    from bonxai_b200 import build as b
    try:
        exe = b.build_tools()
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"could not build tools/cpp/dropin_bench.cpp: {e!r}"}
    out = {}
    with tempfile.NamedTemporaryFile(suffix=".bin", dir="/tmp") as f:
        f.write(np.array([len(scans), len(scans[0][0])], np.int64).tobytes())
        for pts, origin in scans:
            f.write(np.asarray(origin, np.float32).tobytes())
            f.write(np.ascontiguousarray(pts, np.float32).tobytes())
        f.flush()
        for mode in ("vector", "pinned", "publish", "publish_ref"):
            r = subprocess.run([exe, f.name, repr(RES), repr(MAX_RANGE), str(warmup), mode], capture_output=True, text=True, timeout=600)
            if r.returncode != 0:
                out[mode] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    return out


REFERENCE_BUDGET_S = 60.0  # CPU seconds of timed inserts the reference arm may spend (the sample is bounded, the per-step rate is not)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the GPU arm at N > 1 inserts N x denser scans (N x 2048 azimuths) into one sharded map: same scans here.
    # Scans are generated one by one and the timed sample stops after REFERENCE_BUDGET_S of CPU time, so that the
    # arm ends within a few minutes whatever --steps says (one 8 x 131,072-point scan costs the CPU ~0.7 s).
    fleet = args.gpus > 1 and args.workload == "fleet"
    az = AZ * max(1, args.gpus) if not fleet else AZ
    n_pts = BEAMS * az * (args.gpus if fleet else 1)
    lib, kind = load_cpu_oracle()
    m = lib.map(RES)
    secs, n = 0.0, 0
    for i in range(args.warmup + args.steps):
        if fleet:  # one step = the scans of all vehicles, inserted one after the other (the reference has no other way)
            step_s = 0.0
            for sensor in range(args.gpus):
                pts, origin = fleet_scan((i, sensor))
                m.insert(pts, origin, MAX_RANGE)
                step_s += m.last_insert_seconds()
        else:
            pts, origin = synth.lidar_scan(i, beams=BEAMS, azimuths=az)
            m.insert(pts, origin, MAX_RANGE)
            step_s = m.last_insert_seconds()
        if i >= args.warmup:
            secs += step_s
            n += 1
            if secs >= REFERENCE_BUDGET_S:
                break
    pts_s = n * n_pts / secs
    line = {
        "impl": "reference", "metric": "insertPointCloud points/sec", "value": pts_s, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": n, "ms_per_step": 1e3 * secs / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
        "config": {"workload": WORKLOAD if args.gpus <= 1 else sharded_workload_name(args.gpus, args.workload),
                   "points_per_scan": n_pts, "host": "single-threaded reference CPU path"},
        "cpu_baseline": {"value": pts_s, "unit": "points/s", "cores": 1, "kind": kind,
                         "sample": f"scans {args.warmup}..{args.warmup + n - 1} of the same sequence, one map ({n} of the {args.steps} steps: "
                                   f"the timed sample is bounded to {REFERENCE_BUDGET_S:.0f} s of CPU time)"},
        "e2e": {"value": pts_s, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, args.warmup
    total = W + K
    # Multi-GPU (round 1): replicas — every rank maps its own stretch of the street (weak scaling, no exchange).
    first_scan = rank * 5000
    scans = gen_scans(first_scan, total)  # before torch/CUDA: the generator forks worker processes

    import torch
    import torch.distributed as dist
    from bonxai_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bonxai_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        with _stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
    capi.load_library()
    # a real (non-blocking-capable) stream, not the legacy default stream: the library launches its kernels with
    # programmatic dependent launch, which the NULL stream serialises
    stream = torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def settle():
        # destroying a map unmaps ~1 GB of pools; let the driver finish that before the next timed run starts
        torch.cuda.synchronize()
        time.sleep(0.3)

    dev_scans = [torch.from_numpy(p).cuda() for p, _ in scans]  # distinct 2 MiB buffers: > L2 in total for K >= 64

    # ---------------- pass 1: synchronous C-ABI call per scan with per-phase CUDA events (kernel shares, counters) ----
    m = capi.ProbabilisticMap(RES)
    m.set_stream(stream.cuda_stream)
    m.set_profiling(True)
    U = V = E = 0
    phases = {k: 0.0 for k in ("classify", "resolve", "mark", "apply", "total")}
    for i in range(W):
        m.insert(capi.DevPtr(dev_scans[i].data_ptr()), scans[i][1], MAX_RANGE, n=N_PTS, stride_bytes=16)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(W, total):
        m.insert(capi.DevPtr(dev_scans[i].data_ptr()), scans[i][1], MAX_RANGE, n=N_PTS, stride_bytes=16)
        c = m.counters()
        U += c["U"]
        V += c["V"]
        E += c["E"]
        ph = m.phase_times()
        for k in phases:
            phases[k] += ph[k]
    e1.record(stream)
    barrier()
    sync_ms = e0.elapsed_time(e1)
    active = m.active_count()
    sync_digest = m.digest()
    del m

    # ---------------- pass 2 ("value"): pipelined inserts, scans resident in HBM, one sync at the end ----------------
    # Two runs on fresh maps, the faster one is reported (both are kept in the line): the host side of the loop is a
    # Python thread on a shared VM core, and one descheduling of ~10 ms is 1/3 of the whole timed region.
    sampler = ClockSampler(local_rank)
    sampler.start()
    value_runs, launches, enqueue_us, tt, pool = [], 0, 0.0, None, None
    run_digests = [sync_digest]
    for rep in range(2 if world == 1 else 1):
        m = capi.ProbabilisticMap(RES)
        m.set_stream(stream.cuda_stream)
        for i in range(W):
            m.insert_async(capi.DevPtr(dev_scans[i].data_ptr()), scans[i][1], MAX_RANGE, n=N_PTS, stride_bytes=16)
        m.sync()
        barrier()
        launches0 = capi.launch_count()
        e0.record(stream)
        th0 = time.perf_counter()
        for i in range(W, total):
            m.insert_async(capi.DevPtr(dev_scans[i].data_ptr()), scans[i][1], MAX_RANGE, n=N_PTS, stride_bytes=16)
        enq = 1e6 * (time.perf_counter() - th0) / K  # host time to enqueue one scan: must stay below the GPU time
        e1.record(stream)
        m.sync()
        barrier()
        ms_rep = e0.elapsed_time(e1)
        if not value_runs or ms_rep < min(value_runs):
            launches, enqueue_us = capi.launch_count() - launches0, enq
        value_runs.append(ms_rep)
        assert m.active_count() == active, "pipelined and synchronous passes disagree"
        tt = m.totals()
        run_digests.append(m.digest())
        try:  # pool usage at the end of the run: bytes mapped behind the leaf / inner pools vs bytes of live leaves
            gst = m.grid().stats()
            pool = {"leaves_live": int(gst["leaves"]), "leaf_bytes_live": int(gst["leaves"]) * 2304, "mapped_bytes": int(gst["mapped_bytes"]),
                    "mapped_over_live": gst["mapped_bytes"] / max(1, gst["leaves"] * 2304)}
        except Exception as e:  # noqa: BLE001
            pool = {"error": repr(e)}
        del m
        settle()
    clocks = sampler.stop()
    ms = min(value_runs)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---------------- pass 3 ("e2e"): the same pipelined C-ABI call with HOST (pinned) buffers ----------------
    pinned = [torch.from_numpy(p).pin_memory().numpy() for p, _ in scans]
    m2 = capi.ProbabilisticMap(RES)
    for i in range(W):
        m2.insert_async(pinned[i], scans[i][1], MAX_RANGE)
    m2.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(W, total):
        m2.insert_async(pinned[i], scans[i][1], MAX_RANGE)
    enqueue_e2e_us = 1e6 * (time.perf_counter() - t0) / K
    m2.sync()
    e2e_s = time.perf_counter() - t0
    e2e_runs = [e2e_s]
    if world == 1:
        # the host side of this pass (a Python loop on a shared VM core) is noisy: repeat once on a fresh map, keep both
        del m2
        settle()
        m2 = capi.ProbabilisticMap(RES)
        for i in range(W):
            m2.insert_async(pinned[i], scans[i][1], MAX_RANGE)
        m2.sync()
        t0 = time.perf_counter()
        for i in range(W, total):
            m2.insert_async(pinned[i], scans[i][1], MAX_RANGE)
        m2.sync()
        e2e_runs.append(time.perf_counter() - t0)
        e2e_s = min(e2e_runs)
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    assert m2.active_count() == active, "host-buffer and device-buffer arms disagree"
    run_digests.append(m2.digest())
    digests_agree = len(set(run_digests)) == 1
    # synchronous host-buffer call per scan (what the drop-in C++ insertPointCloud does), for reference
    m3 = capi.ProbabilisticMap(RES)
    for i in range(W):
        m3.insert(pinned[i], scans[i][1], MAX_RANGE)
    t0 = time.perf_counter()
    for i in range(W, total):
        m3.insert(pinned[i], scans[i][1], MAX_RANGE)
    e2e_sync_s = time.perf_counter() - t0
    del m2, m3

    tot = torch.tensor([U, V, E, launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    U_all, V_all, E_all, launches_all = [float(x) for x in tot.tolist()]

    if rank == 0:
        peaks, peak_kind = load_peaks()
        secs = ms_max * 1e-3
        pts_s = world * K * N_PTS / secs
        dom = max(("classify", "resolve", "mark", "apply"), key=lambda k: phases[k])
        dom_us = phases[dom] / K  # average CUDA-event duration of the dominant phase per scan (this rank)
        alg_bytes = 16.0 * N_PTS + 8.0 * (U / K)
        achieved = alg_bytes / (dom_us * 1e-6) / 1e9
        peak = float(peaks["hbm_gbs"])
        line = {
            "metric": "insertPointCloud points/sec", "value": pts_s, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int32",
            "data": "synthetic", "value_runs_ms_per_step": [r / K for r in value_runs], "value_reported": "min of the runs",
            "config": {"workload": WORKLOAD, "points_per_scan": N_PTS, "parallelism": "single" if world == 1 else f"replicas x{world}",
                       "call": "bnx_map_insert_async_f32 per scan + one bnx_map_sync (pipelined, no host sync per scan)",
                       "l2": f"{total} distinct 2 MiB scan buffers resident in HBM ({total * 2} MiB: they stay in the 126 MB L2 between the passes when "
                             "the run is short), each read once per pass; the map itself is state carried between scans",
                       "active_cells_end": active},
            "voxel_updates_per_s": U_all / secs, "ray_visits_per_s": V_all / secs, "rays_per_s": E_all / secs,
            "updates_per_scan": U / K, "visits_per_scan": V / K,
            "phase_us_per_scan": {k: phases[k] / K for k in phases},
            "sync_call": {"ms_per_step": sync_ms / K, "points_per_s": K * N_PTS / (sync_ms * 1e-3),
                          "host_buffers_ms_per_step": 1e3 * e2e_sync_s / K, "host_buffers_points_per_s": K * N_PTS / e2e_sync_s,
                          "note": "one synchronous bnx_map_insert_f32 per scan (the drop-in C++ insertPointCloud); value/e2e use the pipelined call"},
            "roofline": {"bound": "hbm", "kernel": {"classify": "k_classify", "resolve": "k_resolve", "mark": "k_mark", "apply": "k_apply_leaves"}[dom],
                         "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(dom), "algorithmic_bytes_per_launch": alg_bytes, "kernel_us": dom_us,
                         "kernel_share_of_step": phases[dom] / max(phases["total"], 1e-9),
                         "step_achieved_gbs": alg_bytes * K / secs / 1e9},
            "e2e": {"value": world * K * N_PTS / e2e_s, "unit": "points/s", "h2d_bytes_per_step": N_PTS * 16, "d2h_bytes_per_step": 64,
                    "ms_per_step": 1e3 * e2e_s / K, "runs_ms_per_step": [1e3 * r / K for r in e2e_runs], "reported": "min of the runs"},
            "gpu_launches": int(launches_all),
            "pool": pool,
            "host_enqueue_us_per_scan": {"device_buffers": enqueue_us, "host_buffers": enqueue_e2e_us},
            "clocks": clocks,
        }
        # CPU baseline beside it: the unmodified reference on one host core, bounded sample of the same sequence
        if world == 1 and not args.no_cpu:
            n_cpu = min(total, args.cpu_scans + 2)
            secs_cpu, n, kind, dig_cpu = time_cpu(scans[:n_cpu], 2)
            # in-run parity: the GPU map after the same scans has the same digest as the CPU reference's dump
            mp_ = capi.ProbabilisticMap(RES)
            for i in range(n_cpu):
                mp_.insert_async(capi.DevPtr(dev_scans[i].data_ptr()), scans[i][1], MAX_RANGE, n=N_PTS, stride_bytes=16)
            mp_.sync()
            dig_gpu = mp_.digest()
            del mp_
            line["parity_in_run"] = bool(dig_gpu == dig_cpu)
            line["parity"] = {"scans": n_cpu, "oracle": kind, "active_cells": dig_gpu[2], "gpu_digest_equals_oracle": dig_gpu == dig_cpu,
                              "passes_agree": bool(digests_agree)}
            if not args.no_dropin:
                dd = run_dropin(scans[:W + min(K, 100)], W)  # bounded: at most 100 timed scans per mode
                line["e2e_dropin"] = {
                    "call": "Bonxai::ProbabilisticMap::insertPointCloud(std::vector<PointXYZ>, origin, max_range) through include/ (C++), "
                            "synchronous per scan, separate process",
                    "pageable_vector": dd.get("vector"), "pinned_allocator_vector": dd.get("pinned"),
                    "ratio_to_pipelined_step": (dd["pinned"]["us_per_scan"] / (1e3 * ms_max / K)) if "us_per_scan" in dd.get("pinned", {}) else None}
                line["e2e_insert_publish"] = {
                    "call": "insertPointCloud + the publisher post-step after EVERY scan (bonxai_server.cpp:176-186,217-251)",
                    "fused_publish_occupied_f32": dd.get("publish"), "reference_style_getOccupiedVoxels_loop": dd.get("publish_ref")}
            line["cpu_baseline"] = {"value": n * N_PTS / secs_cpu, "unit": "points/s", "cores": 1, "kind": kind,
                                    "sample": f"scans 2..{n_cpu - 1} of the same sequence ({n} scans, {secs_cpu:.1f} s), single thread",
                                    "ms_per_step": 1e3 * secs_cpu / n}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _stdout_to_stderr:
    """NCCL prints its version banner on stdout at the first communicator init; the contract is ONE JSON line there."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def _slice_scan(job):
    scan, azimuths, lo, hi = job
    return synth.lidar_scan(scan, beams=BEAMS, azimuths=azimuths, index_range=(lo, hi))


def _full_scan(job):
    scan, azimuths = job
    return synth.lidar_scan(scan, beams=BEAMS, azimuths=azimuths)


FLEET_SPACING = 150.0  # metres between the parallel streets of the fleet workload (> 2 x max_range + margin)


def fleet_scan(job):
    """scan `step` of vehicle `sensor`: the config-#3 generator with its own seed (other buildings), on a street
    FLEET_SPACING * sensor metres to the side; pcl::PointXYZ layout"""
    step, sensor = job
    pts, origin = synth.lidar_scan(step, beams=BEAMS, azimuths=AZ, seed=7 + sensor)
    shift = np.float32([0.0, FLEET_SPACING * sensor, 0.0])
    pts = pts.copy()
    pts[:, :3] += shift
    return pts, origin + shift


def sharded_workload_name(world: int, kind: str) -> str:
    if kind == "fleet":
        return (f"fleet of {world} lidar64x2048 vehicles on parallel streets {FLEET_SPACING:.0f} m apart, one scan per vehicle and step "
                f"({world} x 131072 pts/step, res 0.1 m, max_range 50 m, 1 m/step) into ONE map")
    return f"lidar64x{AZ * world}_seq({BEAMS * AZ * world} pts/scan = {world} x 131072, res 0.1 m, max_range 50 m, 1 m/scan)"


PARITY_SCANS = 8  # scans of the in-run parity check (sharded map == 1-GPU map == CPU oracle, by digest)


def run_gpu_sharded(args):
    """N > 1: ONE map sharded by root key over the N GPUs. Weak scaling: the scan grows with N
    (N x 2048 azimuths -> N x 131,072 points from one origin), every rank holds 131,072 of them; per scan two record
    exchanges + one flag exchange over NVLink (peer-memory stores; BNX_SHARD_EXCHANGE=nccl: NCCL collectives)."""
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, args.warmup
    total = W + K
    fleet = args.workload == "fleet"
    az = AZ * world
    n_scan = BEAMS * az  # points per step over all ranks (fleet: world scans of 131,072 points)
    lo, hi = [(n_scan * r) // world for r in (rank, rank + 1)]
    import multiprocessing as mp
    procs = min(total, max(1, ((os.cpu_count() or 2) - 1) // world), 16)
    n_parity = min(PARITY_SCANS, total)
    # the whole steps of the parity check (rank 0 feeds them to a 1-GPU map and to the CPU oracle, scan after scan)
    if fleet:
        jobs, fn = [(s, rank) for s in range(total)], fleet_scan
        full_jobs, full_fn = [(s, v) for s in range(n_parity) for v in range(world)], fleet_scan
    else:
        jobs, fn = [(s, az, lo, hi) for s in range(total)], _slice_scan
        full_jobs, full_fn = [(s, az) for s in range(n_parity)], _full_scan
    full = []
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            slices = pool.map(fn, jobs, chunksize=max(1, total // (procs * 4)))
            if rank == 0:
                full = pool.map(full_fn, full_jobs)
    else:
        slices = [fn(j) for j in jobs]
        if rank == 0:
            full = [full_fn(j) for j in full_jobs]
    # fleet: the origins of ALL vehicles at every step (the street offsets are exact in float32)
    origins = [np.array([slices[i][1] + np.float32([0.0, FLEET_SPACING * (v - rank), 0.0]) for v in range(world)], np.float64)
               for i in range(total)] if fleet else None

    import torch
    import torch.distributed as dist
    from bonxai_b200 import capi
    from bonxai_b200.sharded import ShardedMap

    torch.cuda.set_device(local_rank % torch.cuda.device_count())
    same_gpu = torch.cuda.device_count() < world  # protocol test on a smaller box: ranks share GPUs, bootstrap over gloo
    with _stdout_to_stderr():
        if same_gpu:
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    bootstrap = "host" if same_gpu else "nccl"
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)  # ShardedMap launches on torch's current stream
    n_local = hi - lo

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def reduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device="cpu" if same_gpu else "cuda")
        dist.all_reduce(t, op=op)
        return [float(x) for x in t.tolist()]

    # NVML is initialised and polled once BEFORE anything is timed: the first queries of a process take milliseconds
    # (more with 8 processes asking at once) and must not fall into the timed region
    sampler = ClockSampler(local_rank % torch.cuda.device_count())
    sampler.start(paused=True)

    dev = [torch.from_numpy(p).cuda() for p, _ in slices]
    n_max = max((n_scan * (r + 1)) // world - (n_scan * r) // world for r in range(world))

    # ---------------- in-run parity: sharded map == 1-GPU map == CPU oracle after the first scans (digests) ------------
    with _stdout_to_stderr():
        sm = ShardedMap(RES, bootstrap=bootstrap)
    def put(m, inp, i):
        if fleet:
            m.insert_fleet(inp, n_local, 16, n_max, origins[i], MAX_RANGE, use_async=True)
        else:
            m.insert(inp, n_local, 16, lo, n_max, slices[i][1], MAX_RANGE, use_async=True)

    for i in range(n_parity):
        put(sm, capi.DevPtr(dev[i].data_ptr()), i)
    sm.sync()
    dig_sharded = sm.digest()  # collective: combined over the ranks
    parity = None
    if rank == 0:
        single = capi.ProbabilisticMap(RES)
        for pts, origin in full:
            single.insert(pts, origin, MAX_RANGE)
        dig_single = single.digest()
        single.close()
        lib, kind = load_cpu_oracle()
        om = lib.map(RES)
        for pts, origin in full:
            om.insert(pts, origin, MAX_RANGE)
        dig_cpu = capi.digest_of_dump(*om.dump())
        del om
        parity = {"scans": n_parity, "oracle": kind, "sharded_equals_single_gpu": dig_sharded == dig_single,
                  "sharded_equals_oracle": dig_sharded == dig_cpu, "active_cells": dig_sharded[2]}
    # maps are closed explicitly, at the same point on every rank (left to the garbage collector, one rank would tear its
    # shard down while another already waits in the next communicator's rendezvous)
    sm.close()
    del sm
    full = None

    def one_run(inputs):
        """fresh sharded map, W warm-up scans, K timed scans; returns device ms (this rank), wall s, stats, digest"""
        with _stdout_to_stderr():
            m = ShardedMap(RES, bootstrap=bootstrap)
        for i in range(W):
            put(m, inputs[i], i)
        m.sync()
        st0, t0s = m.stats(), m.totals()
        barrier()
        l0 = capi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tw = time.perf_counter()
        calls = []
        for i in range(W, total):
            tc = time.perf_counter()
            put(m, inputs[i], i)
            calls.append(time.perf_counter() - tc)
        enq = time.perf_counter() - tw
        e1.record(stream)
        m.sync()
        wall = time.perf_counter() - tw
        barrier()
        st1, t1s = m.stats(), m.totals()
        d = {k: st1[k] - st0[k] for k in ("attempts", "replays", "mailbox_setups", "drains", "sync_retries", "scans")}
        d["leaf_inbox_cap"], d["leaf_inbox_max_fill"] = st1["leaf_inbox_cap"], st1["leaf_inbox_max_fill"]
        return dict(ms=e0.elapsed_time(e1), wall=wall, enqueue_us=1e6 * enq / K, median_call_us=1e6 * sorted(calls)[K // 2], stats=d,
                    launches=capi.launch_count() - l0, totals={k: t1s[k] - t0s[k] for k in t1s}, digest=m.digest(), map=m)

    dev_inputs = [capi.DevPtr(d.data_ptr()) for d in dev]
    sampler.resume()
    runs = []
    for rep in range(2):  # two runs on fresh maps, the faster one is reported (both are kept in the line)
        r = one_run(dev_inputs)
        r["ms_max"] = reduce([r["ms"]], dist.ReduceOp.MAX)[0]
        runs.append(r)
        if rep == 0:
            r.pop("map").close()
    clocks = sampler.stop()
    best = min(runs, key=lambda r: r["ms_max"])
    ms_max = best["ms_max"]
    U, E = best["totals"]["U"], best["totals"]["E"]
    V = best["totals"]["V"] + best["totals"]["N"]  # per-rank V counts ray cells only
    active = runs[1]["digest"][2]
    exchange = runs[1]["map"].exchange_kind()
    # stage times of the sharded scan (CUDA events on the map's stream, one scan at a time; every stage includes the
    # wait for the peers' records it consumes): a separate, untimed pass over the first scans again
    sm = runs[1].pop("map")
    sm.map.set_profiling(True)
    acc, reps = {}, min(20, K)
    for i in range(W, W + reps):
        put(sm, dev_inputs[i], i)
        sm.sync()
        pt = sm.map.phase_times()
        for k, name in (("classify", "begin"), ("resolve", "resolve_mark"), ("mark", "merge"), ("apply", "apply"), ("total", "total"),
                        ("sub6", "resolve_mark.wait_dedupe_resolve"), ("sub7", "resolve_mark.mark")):
            acc[name] = acc.get(name, 0.0) + pt[k] / reps
    sm.map.set_profiling(False)
    sm.close()
    del sm

    # ---------------- e2e: the same call with pinned HOST buffers, wall clock ----------------
    pinned = [torch.from_numpy(p).pin_memory().numpy() for p, _ in slices]
    e2e_runs = []
    for rep in range(2):
        r = one_run(pinned)
        r.pop("map").close()
        r["wall_max"] = reduce([r["wall"]], dist.ReduceOp.MAX)[0]
        e2e_runs.append(r)
    e2e_best = min(e2e_runs, key=lambda r: r["wall_max"])
    e2e_s = e2e_best["wall_max"]

    U_all, V_all, E_all, launches_all = reduce([U, V, E, best["launches"]], dist.ReduceOp.SUM)
    stats_keys = ("attempts", "replays", "mailbox_setups", "drains", "sync_retries", "scans")
    stats_max = dict(zip(stats_keys, reduce([best["stats"][k] for k in stats_keys], dist.ReduceOp.MAX)))
    if rank == 0:
        peaks, peak_kind = load_peaks()
        secs = ms_max * 1e-3
        alg_bytes = 16.0 * n_scan + 8.0 * (U_all / K)  # whole scan, all ranks
        achieved = alg_bytes * K / secs / 1e9
        peak = float(peaks["hbm_gbs"]) * world
        acc["resolve_mark.emit"] = acc["resolve_mark"] - acc["resolve_mark.wait_dedupe_resolve"] - acc["resolve_mark.mark"]
        stage = max(("begin", "resolve_mark.wait_dedupe_resolve", "resolve_mark.mark", "resolve_mark.emit", "merge", "apply"), key=lambda k: acc[k])
        digests_agree = len({tuple(r["digest"]) for r in runs + e2e_runs}) == 1
        parity["runs_agree"] = digests_agree
        parity_ok = bool(parity["sharded_equals_single_gpu"] and parity["sharded_equals_oracle"] and digests_agree)
        line = {
            "metric": "insertPointCloud points/sec", "value": K * n_scan / secs, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int32", "data": "synthetic",
            "value_runs_ms_per_step": [r["ms_max"] / K for r in runs], "value_reported": "min of the runs (max over ranks each)",
            "config": {"workload": sharded_workload_name(world, args.workload),
                       "points_per_scan": n_scan, "points_per_gpu_per_scan": n_local,
                       "parallelism": f"one map sharded by root key over {world} GPUs, pipelined (no host sync per scan); " + (
                           "per scan two record exchanges + one flag exchange as NVLink peer-memory stores from the producing kernels into the owners' "
                           "mailboxes (CUDA IPC), arrival flags instead of collectives; NCCL only bootstraps" if exchange == "p2p" else
                           "per scan 2 all-to-all (grouped ncclSend/Recv) + 1 all-reduce(16 B)"),
                       "exchange": exchange, "ranks_share_gpus": same_gpu,
                       "l2": f"{total} distinct 2 MiB scan slices per GPU resident in HBM (they fit the 126 MB L2 together for short runs); "
                             "the map itself is state carried between scans",
                       "active_cells_end": int(active)},
            "parity_in_run": parity_ok, "parity": parity,
            "pipeline": {**{k: int(v) for k, v in stats_max.items()}, "leaf_inbox_cap": best["stats"]["leaf_inbox_cap"],
                         "leaf_inbox_max_fill": best["stats"]["leaf_inbox_max_fill"],
                         "note": "library counters over the timed scans, max over ranks (bnx_map_shard_stats): a healthy run has attempts == scans == steps "
                                 "and no replay, retry or mailbox re-creation"},
            "voxel_updates_per_s": U_all / secs, "ray_visits_per_s": V_all / secs, "rays_per_s": E_all / secs,
            "updates_per_scan": U_all / K, "visits_per_scan": V_all / K,
            "phase_us_per_scan": {**acc, "limiting_stage": stage,
                                  "note": "rank 0, one scan at a time; begin = classify + bucket, resolve_mark = wait(exchange 1) + dedupe + "
                                          "resolve + mark + emit, merge = wait(exchange 2) + merge + flags, apply = wait(flags) + apply"},
            "roofline": {"bound": "hbm", "kernel": "whole sharded step (per-kernel events are taken in the 1-GPU run)", "achieved": achieved, "peak": peak,
                         "peak_kind": peak_kind + f" x{world}", "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": {"value": K * n_scan / e2e_s, "unit": "points/s", "h2d_bytes_per_step": n_local * 16, "d2h_bytes_per_step": 64,
                    "ms_per_step": 1e3 * e2e_s / K, "runs_ms_per_step": [1e3 * r["wall_max"] / K for r in e2e_runs], "reported": "min of the runs"},
            "gpu_launches": int(launches_all),
            "host_enqueue_us_per_scan": {"device_buffers": best["enqueue_us"], "host_buffers": e2e_best["enqueue_us"],
                                         "median_call_device_buffers": best["median_call_us"], "median_call_host_buffers": e2e_best["median_call_us"],
                                         "note": "rank 0; mean = loop time / scans (includes the collective drain every 64 scans, which waits for the GPUs), "
                                                 "median = one insert call"},
            "clocks": clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()


def main():
    if os.environ.get("BNX_BENCH_WATCHDOG"):  # debugging aid: dump every thread's Python stack and exit if the run hangs
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BNX_BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-scans", type=int, default=60, help="scans timed for the cpu_baseline (about 10-15 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the C++ drop-in arm (e2e_dropin / e2e_insert_publish)")
    ap.add_argument("--workload", default="fleet", choices=["fleet", "dense-scan"],
                    help="N > 1: 'fleet' = N vehicles, one 131,072-point scan each per step, into one sharded map (work grows with N); "
                         "'dense-scan' = ONE sensor whose scan is N x denser (round 1's workload: same voxels, more points)")
    ap.add_argument("--mode", default="sharded", choices=["sharded", "replicas"],
                    help="N > 1: one root-key-sharded map (default) or N independent maps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.mode == "sharded":
        run_gpu_sharded(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
