"""Where does the end-to-end step (setrates! on the host + logpdf!) spend its time beyond the sweep kernel?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import bench
import beluga_b200 as B

a = argparse.Namespace(families=int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, taxa=100, wgds=10, max_count=50)
nw, model, ext, t, l, names, counts = bench.build_workload(a, 0, device="cuda:0")
X, w, m = B.profile(t, l, (names, counts))
data = B.PArray(np.ascontiguousarray(X, dtype=np.int32), w, device=0)
for i in ext:
    B.extend_(data, i)
B.set_(data)
rates = model.getrates()
rng = np.random.default_rng(0)
B.logpdf_(model, data)
T = {"setrates": 0.0, "sync": 0.0, "run": 0.0}
reps = 30
for r in range(reps):
    x = rates * np.exp(rng.normal(0, 1e-3, size=rates.shape))
    t0 = time.perf_counter()
    model.setrates_(x)
    t1 = time.perf_counter()
    data.sync(model); data.synchronize()
    t2 = time.perf_counter()
    data._run(0, 1)
    t3 = time.perf_counter()
    T["setrates"] += t1 - t0; T["sync"] += t2 - t1; T["run"] += t3 - t2
print("patterns", len(data), {k: "%.0f us" % (v / reps * 1e6) for k, v in T.items()})
data.set_timing(True)
data._run(0, 1)
print("device: prune %.0f us, root+reduce %.0f us" % tuple(x * 1e3 for x in data.last_timing()))
