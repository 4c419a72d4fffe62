"""Only the rjMCMC leg of bench.py (config C4), repeated, for A/B runs of host-side changes."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--mcmc-families", type=int, default=100000)
ap.add_argument("--mcmc-generations", type=int, default=40)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
for _ in range(a.reps):
    r = bench.mcmc_leg(a, 0)
    print(json.dumps({k: r[k] for k in ("value", "proposals_per_s", "acceptance")}))
