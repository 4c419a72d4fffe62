"""How much of a warp's work is useful?  Every loop of the sweep kernel runs to the warp's maximum count, so the
cost of a step for a warp is set by its heaviest family.  This script draws the bench workload (C5) on the CPU,
asks the library for its device order (blg_family_order, host-only) and compares, under a simple work model
(contractions S0²/2 + S1²/2, merge S0·S1, plus a fixed per-step overhead), the work the warps do with the work
their families need.  No GPU required."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import beluga_b200 as B
from beluga_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--families", type=int, default=200_000)
ap.add_argument("--taxa", type=int, default=100)
ap.add_argument("--wgds", type=int, default=10)
ap.add_argument("--max-count", type=int, default=50)
ap.add_argument("--overhead", type=float, default=150.0, help="fixed cost of a step, in FMA units")
ap.add_argument("--order", default="lib", choices=["lib", "input", "sum"])
a = ap.parse_args()

cache = "/tmp/blg_c5_X_%d_%d_%d_%d.npy" % (a.families, a.taxa, a.wgds, a.max_count)
if os.path.exists(cache):
    nw, model, ext, t, l, names, counts = bench.build_workload(a, 0, device="cpu", families=8)
    X = np.load(cache)
else:
    nw, model, ext, t, l, names, counts = bench.build_workload(a, 0, device="cpu")
    X, w, m = B.profile(t, l, (names, counts))
    X = np.ascontiguousarray(X, dtype=np.int32)
    np.save(cache, X)
npat, ncols = X.shape
print("patterns", npat, "columns", ncols)
if a.order == "lib":
    perm = np.zeros(npat, dtype=np.int32)
    L = _lib.lib()
    rc = L.blg_family_order(X.ctypes.data_as(C.POINTER(C.c_int32)), ncols, npat, perm.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
elif a.order == "sum":
    perm = np.argsort(-X[:, 0], kind="stable").astype(np.int32)
else:
    perm = np.arange(npat, dtype=np.int32)
Xd = X[perm]
pad = (-npat) % 32
if pad:
    Xd = np.vstack([Xd, np.zeros((pad, ncols), dtype=np.int32)])
Xw = Xd.reshape(-1, 32, ncols)
# steps: internal nodes of the species tree (WGD nodes add single-child steps on aliased counts: ignored here)
def walk(n, out):
    for c in n.c:
        walk(c, out)
    if n.c:
        out.append(n)
    return out
steps = [n for n in walk(t, []) if len(n.c) == 2]
need = 0.0
done = 0.0
fma_need = fma_done = 0.0
for n in steps:
    c0, c1 = n.c[0].i - 1, n.c[1].i - 1
    leaf0, leaf1 = not n.c[0].c, not n.c[1].c
    def cost(s0, s1):
        s0 = s0 + 1.0
        s1 = s1 + 1.0
        return (0 if leaf0 else 0.5 * s0 * s0) + (0 if leaf1 else 0.5 * s1 * s1) + s0 * s1 + (s0 if leaf0 else 0) + (s1 if leaf1 else 0)
    f_need = cost(Xw[:, :, c0].astype(np.float64), Xw[:, :, c1].astype(np.float64)).sum() / 32.0
    f_done = cost(Xw[:, :, c0].max(axis=1).astype(np.float64), Xw[:, :, c1].max(axis=1).astype(np.float64)).sum()
    fma_need += f_need
    fma_done += f_done
nw_ = Xw.shape[0]
ov = a.overhead * len(steps) * nw_
print("warps %d  steps %d" % (nw_, len(steps)))
print("FMA work: needed %.3g  issued %.3g  efficiency %.3f" % (fma_need, fma_done, fma_need / fma_done))
print("with %.0f-FMA fixed overhead/step: efficiency %.3f ; overhead share of issued %.3f" % (a.overhead, (fma_need + ov) / (fma_done + ov), ov / (fma_done + ov)))
