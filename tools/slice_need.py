"""How many doubles per lane does each warp's heaviest step need in the shared slice (sum of the children's warp-max
counts + k + 1)?  Answers what a smaller slice (→ more blocks per SM) would send to the global-scratch path.
CPU only; uses the cached C5 pattern matrix of tools/warp_efficiency.py when present."""
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
a = ap.parse_args()
cache = "/tmp/blg_c5_X_%d_%d_%d_%d.npy" % (a.families, a.taxa, a.wgds, a.max_count)
nw, model, ext, t, l, names, counts = bench.build_workload(a, 0, device=None, families=8 if os.path.exists(cache) else None)
if os.path.exists(cache):
    X = np.load(cache)
else:
    X, w, m = B.profile(t, l, (names, counts))
    X = np.ascontiguousarray(X, dtype=np.int32)
    np.save(cache, X)
npat, ncols = X.shape
perm = np.zeros(npat, dtype=np.int32)
rc = _lib.lib().blg_family_order(X.ctypes.data_as(C.POINTER(C.c_int32)), ncols, npat, perm.ctypes.data_as(C.POINTER(C.c_int32)))
assert rc == 0
Xd = X[perm]
pad = (-npat) % 32
if pad:
    Xd = np.vstack([Xd, np.zeros((pad, ncols), dtype=np.int32)])
Xw = Xd.reshape(-1, 32, ncols).max(axis=1)


def walk(n, out):
    for c in n.c:
        walk(c, out)
    if n.c:
        out.append(n)
    return out


need = np.zeros(Xw.shape[0], dtype=np.int64)
steps_over = {s: 0 for s in (32, 36, 40, 42, 46, 50, 54, 58, 62, 70)}
nsteps = 0
for n in walk(t, []):
    k = len(n.c)
    nd = sum(Xw[:, c.i - 1] for c in n.c) + k + 1
    busy = Xw[:, n.i - 1] > 0
    nsteps += int(busy.sum())
    for s in steps_over:
        steps_over[s] += int((nd[busy] > s).sum())
    need = np.maximum(need, nd)
print("patterns %d  warps %d  busy (warp, step) pairs %d" % (npat, Xw.shape[0], nsteps))
for s in sorted(steps_over):
    print("slice %2d doubles: %5.2f %% of the warps have a step above it, %6.3f %% of the busy steps"
          % (s, 100.0 * (need > s).mean(), 100.0 * steps_over[s] / max(nsteps, 1)))
