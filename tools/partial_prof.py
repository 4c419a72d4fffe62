"""Per-step warp-cycle profile (BLG_PROF=2) of one node-local recompute on the C4 workload."""
import os, sys
os.environ["BLG_PROF"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import beluga_b200 as B
from beluga_b200.model import build_model, update_
from beluga_b200.sim import rand_profiles_torch

nw = open(os.path.join(os.path.dirname(__file__), "..", "tests", "data", "12monocots.nw")).readline().strip()
m0, _, _ = build_model(nw, 0.004, 0.003, 0.9, B.BRANCH)
names, counts = rand_profiles_torch(m0, 100000, seed=1, device="cuda:0", max_rootsum=60, chunk=1 << 19)
model, data = B.DLWGD(nw, (names, counts), 0.004, 0.003, 0.9, B.BRANCH, device=0)
sys.stderr.write("=== full sweep\n")
B.logpdf_(model, data); B.set_(data)
leaf = [n for n in model.nodes.values() if not n.c][3]
sys.stderr.write("=== partial from leaf %d\n" % leaf.i)
update_(leaf, λ=leaf["λ"] * 1.01, μ=leaf["μ"])
B.logpdf_(leaf, data)
