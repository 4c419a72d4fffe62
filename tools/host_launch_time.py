"""Host time of ENQUEUEING one resident likelihood call (blg_logpdf_full_async) against its device time, launch by
launch (BLG_GRAPH=0) and replayed from the captured CUDA graph (default).  C5 workload, 1 GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import ctypes as C
import numpy as np
import torch
import bench
import beluga_b200 as B

fam = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
a = argparse.Namespace(families=fam, taxa=100, wgds=10, max_count=50)
nw, model, ext, t, l, names, counts = bench.build_workload(a, 0, device="cuda:0")
X, w, m = B.profile(t, l, (names, counts))
X = np.ascontiguousarray(X, dtype=np.int32)
for graph in ("0", "1"):
    os.environ["BLG_GRAPH"] = graph
    data = B.PArray(X, w, device=0, distributed=True)       # (own torch stream, async entries)
    for i in ext:
        B.extend_(data, i)
    B.set_(data)
    ll = B.logpdf_(model, data)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    ptr = C.c_void_p(out.data_ptr())
    st = data.torch_stream
    for _ in range(5):
        data._ck(data.L.blg_logpdf_full_async(data.h, ptr))
    torch.cuda.synchronize()
    reps = 50
    # (a) host time of a call when the queue is empty: sync before every call
    th = 0.0
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data._ck(data.L.blg_logpdf_full_async(data.h, ptr))
        th += time.perf_counter() - t0
    torch.cuda.synchronize()
    # (b) back to back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    t0 = time.perf_counter()
    for _ in range(reps):
        data._ck(data.L.blg_logpdf_full_async(data.h, ptr))
    tq = time.perf_counter() - t0
    e1.record(st)
    torch.cuda.synchronize()
    print("BLG_GRAPH=%s patterns %d: host enqueue %.0f us/call (idle queue), %.0f us/call back to back; device %.1f us/step; "
          "logpdf %.6f (%.6f); graphs %s" % (graph, len(data), th / reps * 1e6, tq / reps * 1e6, e0.elapsed_time(e1) / reps * 1e3,
                                              float(out.item()), ll, data.graph_stats()), flush=True)
    del data
