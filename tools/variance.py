"""Run-to-run spread of the pipelined insert on the LiDAR bench workload: R fresh maps x K scans each."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bonxai_b200 import capi
import torch

K, W, R = 300, 10, 5
scans = bench.gen_scans(0, K + W)
dev = [torch.from_numpy(p).cuda() for p, _ in scans]
pinned = [torch.from_numpy(p).pin_memory().numpy() for p, _ in scans]
stream = torch.cuda.Stream()
for mode in ("device", "host"):
    for r in range(R):
        m = capi.ProbabilisticMap(bench.RES)
        m.set_stream(stream.cuda_stream)
        for i in range(W):
            src = capi.DevPtr(dev[i].data_ptr()) if mode == "device" else pinned[i]
            m.insert_async(src, scans[i][1], bench.MAX_RANGE, n=bench.N_PTS, stride_bytes=16)
        m.sync()
        cap0 = m.grid().stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        marks = []
        for i in range(W, W + K):
            src = capi.DevPtr(dev[i].data_ptr()) if mode == "device" else pinned[i]
            ta = time.perf_counter()
            m.insert_async(src, scans[i][1], bench.MAX_RANGE, n=bench.N_PTS, stride_bytes=16)
            dt = time.perf_counter() - ta
            if dt > 1e-3:
                marks.append((i, round(dt * 1e3, 2)))
        m.sync()
        dt = time.perf_counter() - t0
        print(mode, r, "us/scan %.1f" % (1e6 * dt / K), "slow enqueues (scan, ms):", marks[:12], cap0, m.grid().stats(), flush=True)
        del m
