import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import beluga_b200 as B
from beluga_b200.model import build_model
from beluga_b200.sim import rand_profiles_torch
nw = open(os.path.join(os.path.dirname(__file__), "..", "tests", "data", "12monocots.nw")).readline().strip()
m0, _, _ = build_model(nw, 0.004, 0.003, 0.9, B.BRANCH)
names, counts = rand_profiles_torch(m0, 100000, seed=1, device="cuda:0", max_rootsum=60, chunk=1 << 19)
print("root-sum percentiles", np.percentile(counts.sum(1), [10, 50, 90, 99]), "leaf max", counts.max())
model, data = B.DLWGD(nw, (names, counts), 0.004, 0.003, 0.9, B.BRANCH, device=0)
B.logpdf_(model, data)
data.set_timing(True)
for _ in range(3):
    B.logpdf_(model, data)
print("full sweep ms (prune, root):", data.last_timing(), "patterns", len(data))
data.set_timing(True)
from beluga_b200.model import update_
for n in sorted(model.nodes.values(), key=lambda x: x.i):
    if n.p is None: continue
    update_(n, λ=n["λ"] * 1.0001, μ=n["μ"]); data.sync(model); data.synchronize()
    data._run(1, n.i)
    t0 = time.perf_counter()
    for _ in range(10): data._run(1, n.i)
    dt = (time.perf_counter() - t0) / 10
    depth = 0; x = n
    while x is not None: depth += 1 if x.c else 0; x = x.p
    print("node %2d internal-path %d: host %.0f us, device prune %.0f us root %.0f us, kernel v%d" % (n.i, depth, dt * 1e6, data.last_timing()[0] * 1e3, data.last_timing()[1] * 1e3, 0))
