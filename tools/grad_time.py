"""gradient(model, data) on the benchmark workload (C5) under several settings of the library's switches.
GRAD_SPECS="BLG_GRAD_PAT=1;BLG_GRAD_PAT=0" (default): the node-pattern adjoint against the per-family adjoint."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import beluga_b200 as B
args = bench.parse()
nw, model, ext, t, l, names, counts = bench.build_workload(args, 0, device="cuda:0")
X, w, m = B.profile(t, l, (names, counts))
out = []
for spec in os.environ.get("GRAD_SPECS", "BLG_GRAD_PAT=1;BLG_GRAD_PAT=0").split(";"):
    kv = [x.split("=") for x in spec.split(",") if x]
    for k, v in kv:
        os.environ[k] = v                       # read when the handle is created
    data = B.PArray(np.ascontiguousarray(X, dtype=np.int32), w, device=0)
    for i in ext:
        B.extend_(data, i)
    B.set_(data)
    B.logpdf_(model, data)
    t0 = time.perf_counter()
    g, ll = data.gradient(model)
    first = (time.perf_counter() - t0) * 1e3
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        g, ll = data.gradient(model)
        ts.append((time.perf_counter() - t0) * 1e3)
    out.append(np.array(g))
    print(f"{spec}: first {first:.1f} ms, then {min(ts):.2f} ms (median {sorted(ts)[3]:.2f}), P = {len(g)}, logL = {ll:.6f}", file=sys.stderr)
    for k, v in kv:
        os.environ.pop(k, None)
    del data
b = out[-1]
for i, a in enumerate(out[:-1]):
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    print(f"setting {i} against the last: max relative difference {rel.max():.3e}", file=sys.stderr)
    for j in np.argsort(-rel)[:4]:
        print(f"    entry {j}: {a[j]:.12e} against {b[j]:.12e} (largest |g| {np.abs(b).max():.3e})", file=sys.stderr)
