"""Host-side cost of getting 10^6 families onto the device: profile() (site patterns, numpy), blg_set_profiles
(family order, upload), node-pattern construction (first likelihood call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import beluga_b200 as B
args = bench.parse()
nw, model, ext, t, l, names, counts = bench.build_workload(args, 0, device="cuda:0")
t0 = time.perf_counter()
X, w, m = B.profile(t, l, (names, counts))
t1 = time.perf_counter()
X = np.ascontiguousarray(X, dtype=np.int32)
data = B.PArray(X, w, device=0)
t2 = time.perf_counter()
for i in ext:
    B.extend_(data, i)
B.set_(data)
ll = B.logpdf_(model, data)
t3 = time.perf_counter()
ll = B.logpdf_(model, data)
t4 = time.perf_counter()
print("profile() %.2f s | PArray (set_profiles) %.2f s | first logpdf (node patterns + links) %.2f s | second logpdf %.4f s | %d patterns" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, X.shape[0]))
