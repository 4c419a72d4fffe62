"""Per-launch timing of the pattern path on the benchmark workload (BLG_PDEBUG=1)."""
import os, sys
os.environ["BLG_PDEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import beluga_b200 as B
args = bench.parse()
nw, model, ext, t, l, names, counts = bench.build_workload(args, 0, device="cuda:0")
X, w, m = B.profile(t, l, (names, counts))
data = B.PArray(np.ascontiguousarray(X, dtype=np.int32), w, device=0)
for i in ext:
    B.extend_(data, i)
B.set_(data)
os.environ["BLG_PDEBUG"] = "0"
print(B.logpdf_(model, data), file=sys.stderr)
data.L.blg_sync(data.h)
data2 = None
print("---- second call ----", file=sys.stderr)
# handle reads BLG_PDEBUG at create: make a new one for the timed pass
os.environ["BLG_PDEBUG"] = "1"
data2 = B.PArray(np.ascontiguousarray(X, dtype=np.int32), w, device=0)
for i in ext:
    B.extend_(data2, i)
B.set_(data2)
B.logpdf_(model, data2)
print("---- warm ----", file=sys.stderr)
B.logpdf_(model, data2)
print(data2.pattern_stats(), data2.last_algorithmic(), file=sys.stderr)
