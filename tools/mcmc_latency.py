"""Where does the time of one MCMC proposal go?  (host python / parameter upload + eps/W build / sweep + root)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import beluga_b200 as B
from beluga_b200.model import build_model, update_
from beluga_b200.sim import rand_profiles_device

nw = open(os.path.join(os.path.dirname(__file__), "..", "tests", "data", "12monocots.nw")).readline().strip()
m0, _, _ = build_model(nw, 0.004, 0.003, 0.9, B.BRANCH)
names, counts = rand_profiles_device(m0, int(sys.argv[1]) if len(sys.argv) > 1 else 100000, seed=1, device=0, max_rootsum=60)
model, data = B.DLWGD(nw, (names, counts), 0.004, 0.003, 0.9, B.BRANCH, device=0)
print("patterns", len(data))
B.logpdf_(model, data); B.set_(data)
rng = np.random.default_rng(0)
T = {"update": 0.0, "sync": 0.0, "logpdf": 0.0, "commit": 0.0}
nodes = [n for n in model.nodes.values() if n.p is not None]
reps = 300
for r in range(reps):
    n = nodes[r % len(nodes)]
    t0 = time.perf_counter()
    update_(n, λ=n["λ"] * float(np.exp(rng.normal(0, 0.01))), μ=n["μ"])
    t1 = time.perf_counter()
    data.sync(model); data.synchronize()
    t2 = time.perf_counter()
    l = data._run(1, n.i)
    t3 = time.perf_counter()
    (data.set_ if r % 2 else data.rev_)()
    t4 = time.perf_counter()
    T["update"] += t1 - t0; T["sync"] += t2 - t1; T["logpdf"] += t3 - t2; T["commit"] += t4 - t3
print({k: "%.1f us" % (v / reps * 1e6) for k, v in T.items()}, "total %.1f us" % (sum(T.values()) / reps * 1e6))
# pure device time of the partial sweeps
data.set_timing(True)
tt = []
for r in range(50):
    n = nodes[r % len(nodes)]
    update_(n, λ=n["λ"] * 1.001, μ=n["μ"]); data.sync(model)
    data._run(1, n.i); tt.append(data.last_timing())
print("device: prune %.1f us root+reduce %.1f us (mean over paths)" % (np.mean([a for a, b in tt]) * 1e3, np.mean([b for a, b in tt]) * 1e3))
