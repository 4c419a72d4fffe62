"""Per-launch timing of the node-pattern gradient sweep on the benchmark workload (BLG_PDEBUG=1, events between launches)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import beluga_b200 as B
args = bench.parse()
nw, model, ext, t, l, names, counts = bench.build_workload(args, 0, device="cuda:0")
X, w, m = B.profile(t, l, (names, counts))
os.environ["BLG_PDEBUG"] = "1"
data = B.PArray(np.ascontiguousarray(X, dtype=np.int32), w, device=0)
for i in ext:
    B.extend_(data, i)
B.set_(data)
devnull = os.open(os.devnull, os.O_WRONLY)
saved = os.dup(2)
os.dup2(devnull, 2)               # warm-up calls: not printed
B.logpdf_(model, data)
data.gradient(model)
data.gradient(model)
os.dup2(saved, 2)
t0 = time.perf_counter()
g, ll = data.gradient(model)
print(f"gradient: {(time.perf_counter() - t0) * 1e3:.2f} ms (with the debug events)", file=sys.stderr)
