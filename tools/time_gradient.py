import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import beluga_b200 as B
from beluga_b200.model import build_model
from beluga_b200.sim import add_random_wgds, rand_profiles_torch, random_tree_newick

def run(name, nw, nfam, lam, mu, eta, wgds, maxsum):
    m0, t, l = build_model(nw, lam, mu, eta, B.BRANCH)
    ext = add_random_wgds(m0, wgds, seed=7) if wgds else []
    names, counts = rand_profiles_torch(m0, nfam, seed=1, device="cuda:0", max_rootsum=maxsum, chunk=1 << 19)
    mb, tb, lb = build_model(nw, lam, mu, eta, B.BRANCH)
    X, w, _ = B.profile(tb, lb, (names, counts))
    data = B.PArray(X, w, device=0)
    for i in ext: B.extend_(data, i)
    B.set_(data)
    t0 = time.perf_counter(); l0 = B.logpdf_(m0, data); data.synchronize(); t1 = time.perf_counter()
    l0 = B.logpdf_(m0, data); data.synchronize(); t2 = time.perf_counter()
    g, ll = data.gradient(m0); t3 = time.perf_counter()
    g, ll = data.gradient(m0); t4 = time.perf_counter()
    print("%s: %d patterns, %d nodes, P=%d: logpdf %.2f ms, gradient %.1f ms (%.0fx a sweep), |g|max %.3g" %
          (name, len(data), len(m0), len(g), (t2 - t1) * 1e3, (t4 - t3) * 1e3, (t4 - t3) / (t2 - t1), np.abs(g).max()))

here = os.path.join(os.path.dirname(__file__), "..", "tests", "data")
run("9dicots", open(os.path.join(here, "9dicots.nw")).readline().strip(), 20000, 1.0, 0.92, 0.66, 3, 40)
run("12monocots", open(os.path.join(here, "12monocots.nw")).readline().strip(), 20000, 0.004, 0.003, 0.9, 0, 60)
run("C5-small", random_tree_newick(100), int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 1.0, 1.2, 0.8, 10, 50)
