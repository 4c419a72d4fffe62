"""Host-side cost of getting 10^6 families onto the device: profile() (site patterns on the device, extended counts in
numpy), blg_set_profiles_columns (bounds, 16-bit staging, upload), node-pattern construction (first likelihood call).
BLG_PDEBUG=1 adds the library's own phase timers."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import beluga_b200 as B
from beluga_b200.profile import site_patterns
args = bench.parse()
nw, model, ext, t, l, names, counts = bench.build_workload(args, 0, device="cuda:0")
torch.cuda.synchronize()
print("count table: %s %s" % (counts.shape, counts.dtype))
t0 = time.perf_counter()
u, mlt = site_patterns(counts, 0)
ta = time.perf_counter()
u, mlt = site_patterns(counts, 0)
tb = time.perf_counter()
print("site_patterns: first call %.2f s, second call %.2f s (%d patterns)" % (ta - t0, tb - ta, u.shape[0]))
t0 = time.perf_counter()
X, w, m = B.profile(t, l, (names, counts))
t1 = time.perf_counter()
data = B.PArray(X, w, device=0)
t2 = time.perf_counter()
for i in ext:
    B.extend_(data, i)
B.set_(data)
t2b = time.perf_counter()
ll = B.logpdf_(model, data)
t3 = time.perf_counter()
ll = B.logpdf_(model, data)
t4 = time.perf_counter()
print("(column-major profile: %s) profile() %.2f s | PArray (set_profiles) %.2f s | extend!/set! %.3f s | first logpdf (ε/W tables, node patterns, links, slabs) %.2f s | "
      "second logpdf %.4f s | %d patterns" % (bool(X.flags.f_contiguous), t1 - t0, t2 - t1, t2b - t2, t3 - t2b, t4 - t3, X.shape[0]))
