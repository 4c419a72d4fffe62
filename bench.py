#!/usr/bin/env python
"""Headline benchmark: family-likelihood evaluations per second of `logpdf!(model, profile)`.

Workload (BASELINE.json configs[4] / SURVEY.md §8d C5): synthetic ultrametric 100-taxon tree
(Kingman coalescent, height 1, seed 20261019) with 10 WGDs, Branch model λ=1.0 μ=1.2 η=0.8,
10^6 families drawn from the model, conditioned on both root clades, site-pattern compressed like
profile() does.  "max count 50" has two readings; both are measured and labelled:
  * headline (`value`): ROOT-SUM <= 50, i.e. m = 51 — the reference's m is the maximum of the extended
    profile, which is the root count (src/model.jl:25,48);
  * `c5_leaf50`: only families with a LEAF count above 50 are rejected, m is data-determined (root
    sums in the hundreds to thousands) — a smaller draw, most of it on the general kernel.
A "step" is one full leaves-to-root sweep over every pattern of the rank's shard.  One process per
GPU; each rank draws its own 10^6 families (weak scaling); the only cross-rank step is the all-reduce
of the scalar, issued by the library on its own NCCL communicator.  With more than one GPU a
`strong_scaling` block is added: ONE 10^6-family draw, compressed globally, sharded contiguously.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # the CUDA path
  python bench.py --impl reference [...]                         # the reference algorithm on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "family-likelihood evals/sec (logpdf!)"
UNIT = "families/s"
SEED = 20261019


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--families", type=int, default=1_000_000, help="families drawn per GPU")
    ap.add_argument("--taxa", type=int, default=100)
    ap.add_argument("--wgds", type=int, default=10)
    ap.add_argument("--max-count", type=int, default=50)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mcmc", action="store_true", help="skip the rjMCMC iterations/s leg")
    ap.add_argument("--mcmc-families", type=int, default=100_000)
    ap.add_argument("--mcmc-generations", type=int, default=40)
    ap.add_argument("--leaf50-families", type=int, default=100_000, help="families of the 'leaf count <= 50' reading (0: skip)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(args, rank, device=None, families=None):
    """-> (newick, model with WGDs, ext ids, names, counts) for this rank."""
    import beluga_b200 as B
    from beluga_b200.model import build_model
    from beluga_b200.sim import add_random_wgds, rand_profiles, rand_profiles_device, random_tree_newick

    nw = random_tree_newick(args.taxa, seed=SEED)
    model, t, l = build_model(nw, 1.0, 1.2, 0.8, B.BRANCH)
    ext = add_random_wgds(model, args.wgds, seed=SEED)
    nfam = families if families is not None else args.families
    if device is not None:         # rand(model, N, condition = rootclades) by the library's own sampler kernel (csrc/sampler.cuh)
        names, counts = rand_profiles_device(model, nfam, seed=SEED + rank, device=int(str(device).split(":")[-1]), max_rootsum=args.max_count)
    else:
        names, counts = rand_profiles(model, nfam, seed=SEED + rank, max_rootsum=args.max_count)
    return nw, model, ext, t, l, names, counts


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons sampled while the timed region runs: NVML in-process (nvidia_ml_py — the same counters
    nvidia-smi prints; no fork of this process, no second NVML client per sample), `nvidia-smi` itself as the fallback.
    With several ranks only rank 0 samples: eight nvidia-smi clients polling the driver stall the other ranks' launches."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, enabled=True):
        super().__init__(daemon=True)
        self.index = index
        self.enabled = enabled
        self.stop_ = threading.Event()
        self.rows = []
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        # LOCAL_RANK counts the visible devices: translate through CUDA_VISIBLE_DEVICES when it lists indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and all(v.isdigit() for v in ids) and self.index < len(ids):
            idx = int(ids[self.index])
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def run(self):
        if not self.enabled:
            return
        try:
            nv, h = self._nvml_handle()
            self.source = "nvml"
            bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                    nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_.is_set():
                try:
                    sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits])
                except Exception:
                    pass
                self.stop_.wait(0.005)         # (in-process query, ~10 us: the timed regions are tens of milliseconds long)
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self.stop_.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop_.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


def sample_pattern_rows(args, device=None, nsample=20000):
    """Leaf-count rows of `nsample` patterns drawn uniformly (seeded) from the compressed 10^6-family set of rank 0 — the
    same pattern mix the CUDA arm evaluates.  Without a GPU to draw the 10^6 families with, a 4000-family CPU draw."""
    import beluga_b200 as B
    if device is None:
        nw, model, ext, t, l, names, counts = build_workload(args, 0, device=None, families=4000)
        return nw, names, counts, "4000-family CPU draw (no GPU to draw the full set)", False
    nw, model, ext, t, l, names, counts = build_workload(args, 0, device=device)
    X, w, m = B.profile(t, l, (names, counts))
    leaf_ids = sorted(l)
    rng = np.random.default_rng(SEED)
    sel = np.sort(rng.choice(X.shape[0], size=min(nsample, X.shape[0]), replace=False))
    rows = X[sel][:, [i - 1 for i in leaf_ids]].astype(np.int64)
    sample_pattern_rows.last = (int(X.shape[0]), int(w.sum()))
    return nw, [l[i] for i in leaf_ids], rows, "uniform sample of the %d patterns of the compressed %d-family set" % (X.shape[0], args.families), True


def cpu_baseline(args, threads, target_seconds, steps=None, warmup=1, sample=None):
    """The reference algorithm (oracle/, literal log-space restatement, built -O3) on host cores, on a bounded
    sample of the same workload.  Test infrastructure used as a yard-stick, never as the product."""
    from oracle.oracle import BRANCH as OBRANCH, OracleDLWGD, max_threads

    # all the host cores this process may use — not OMP_NUM_THREADS, which torchrun pins to 1
    threads = threads or max(max_threads(), len(os.sched_getaffinity(0)))
    if sample is None:
        sample = sample_pattern_rows(args, None)
    nw, names, counts, what, same_mix = sample

    # the oracle API inserts a WGD at distance t above a node; replay in insertion order
    def replay(o):
        import beluga_b200 as B
        from beluga_b200.model import build_model
        m2, _, _ = build_model(nw, 1.0, 1.2, 0.8, B.BRANCH)
        rng = np.random.default_rng(SEED)
        for _ in range(args.wgds):
            ids = sorted(m2.nodes)
            w = np.array([m2.nodes[i]["t"] for i in ids], dtype=np.float64)
            i = int(rng.choice(ids, p=w / w.sum()))
            n = m2.nodes[i]
            tt = rng.random() * n["t"]
            if not (n["t"] - tt > 0.0):
                tt = 0.5 * n["t"]
            q = float(rng.beta(1.0, 3.0))
            m2.addwgd_(n, tt, q)
            o.addwgd(i, tt, q)
            o.extend(i)

    n = 64
    o = OracleDLWGD(nw, names, counts[:n], 1.0, 1.2, 0.8, OBRANCH, threads=threads)
    replay(o)
    t0 = time.perf_counter()
    o.logpdf()
    dt = time.perf_counter() - t0
    per = dt / max(o.npatterns, 1)
    reps = steps if steps else 5
    n = int(min(len(counts), max(64, target_seconds / max(per, 1e-9) / (reps + warmup))))
    # (rows are in random order already when they come from the GPU arm's pattern set: a prefix is a uniform sample)
    o = OracleDLWGD(nw, names, counts[:n], 1.0, 1.2, 0.8, OBRANCH, threads=threads)
    replay(o)
    for _ in range(warmup):
        o.logpdf()
    t0 = time.perf_counter()
    for _ in range(reps):
        val = o.logpdf()
    dt = (time.perf_counter() - t0) / reps
    return {"value": o.npatterns / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d patterns: %s; %d full sweeps; literal log-space C++ restatement of src/model.jl:402-463 (gcc -O3, OpenMP over "
                      "families); Julia is not installed in this image" % (o.npatterns, what, reps),
            "same_pattern_mix": same_mix, "ms_per_step": dt * 1e3, "logpdf": val}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.oracle import max_threads
    threads = max(max_threads(), len(os.sched_getaffinity(0)))
    sample = None
    try:                                   # the same pattern mix as the CUDA arm when there is a GPU to draw the 10^6 families with
        import torch
        if torch.cuda.is_available():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            sample = sample_pattern_rows(args, "cuda:%d" % local)
            rng = np.random.default_rng(SEED + 1)
            sample = (sample[0], sample[1], sample[2][rng.permutation(sample[2].shape[0])], sample[3], sample[4])
    except Exception:
        sample = None
    cb = cpu_baseline(args, threads, target_seconds=max(30.0, args.cpu_seconds), steps=args.steps, warmup=args.warmup, sample=sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, *getattr(sample_pattern_rows, "last", (None, None))),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "same_pattern_mix")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def mcmc_leg(args, local, distributed=False):
    """rjMCMC iterations/s (second half of BASELINE.json's metric), config C4: 12monocots tree, 10^5 synthetic
    families from the model, one GPU, `rjmcmc!` generations (one add/remove jump + one MWG sweep each) replayed
    through the public API by beluga_b200.mcmc (flat prior, fixed-width proposals)."""
    import beluga_b200 as B
    from beluga_b200.mcmc import RevJumpChain
    from beluga_b200.model import build_model
    from beluga_b200.sim import rand_profiles_device

    nw = open(os.path.join(ROOT, "tests", "data", "12monocots.nw")).readline().strip()
    m0, _, _ = build_model(nw, 0.004, 0.003, 0.9, B.BRANCH)
    names, counts = rand_profiles_device(m0, args.mcmc_families, seed=SEED, device=local, max_rootsum=60)
    # distributed: every rank draws the same table (same seed), the patterns are compressed globally and sharded contiguously;
    # every rank runs the same chain (same seed, same all-reduced log-likelihoods, hence the same decisions)
    model, data = B.DLWGD(nw, (names, counts), 0.004, 0.003, 0.9, B.BRANCH, device=local, distributed=distributed)
    from beluga_b200.priors import IRRevJumpPrior
    # C2/C4: the chain runs under IRRevJumpPrior(Tl = treelength(model)) (README.md:75-88 of the reference)
    chain = RevJumpChain(data, model, seed=SEED, step=0.02, prior=IRRevJumpPrior(Tl=model.treelength())).init_()
    chain.rjmcmc_(5, trace=0)                      # warm-up
    data.synchronize()
    e0, p0 = chain.evals, chain.proposed
    t0 = time.perf_counter()
    chain.rjmcmc_(args.mcmc_generations, trace=0)
    data.synchronize()
    dt = time.perf_counter() - t0
    try:                                           # BASELINE.json's "gradient-based" wording: one gradient(model, data) on the chain's state
        if distributed:
            raise RuntimeError("skipped on sharded data")
        data.gradient(model)
        t1 = time.perf_counter()
        gg, _ = data.gradient(model)
        grad_ms, grad_p = (time.perf_counter() - t1) * 1e3, int(len(gg))
    except Exception:
        grad_ms, grad_p = None, None
    return {"metric": "rjMCMC iters/sec", "value": args.mcmc_generations / dt, "unit": "generations/s", "gradient_ms": grad_ms, "gradient_P": grad_p,
            "config": {"workload": "C4: 12monocots, %d synthetic families (%d patterns on this rank), rjmcmc! generations, %s" % (
                args.mcmc_families, len(data), "families sharded over the ranks, every proposal all-reduced" if distributed else "1 GPU"),
                       "nodes": len(model), "wgds_at_end": model.nwgd()},
            "likelihood_calls_per_s": (chain.evals - e0) / dt, "proposals_per_s": (chain.proposed - p0) / dt,
            "acceptance": chain.accepted / max(chain.proposed, 1), "generations": args.mcmc_generations}


def c2_leg(args, local):
    """Config C2 (BASELINE.json configs[1]): the 12monocots tree with its shipped profile CSV (1000 families), a
    fixed-dimension `mcmc!` chain under IRRevJumpPrior(Tl = treelength(model)) (README.md:75-88 of the reference), replayed
    through the public API: generations/s (one generation = one Metropolis-within-Gibbs sweep over the 23 nodes)."""
    import beluga_b200 as B
    from beluga_b200.mcmc import RevJumpChain
    from beluga_b200.priors import IRRevJumpPrior
    nw = open(os.path.join(ROOT, "tests", "data", "12monocots.nw")).readline().strip()
    path = os.path.join(ROOT, "tests", "data", "12monocots-f01-1000.csv")
    names = open(path).readline().strip().split(",")
    counts = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
    model, data = B.DLWGD(nw, (names, counts), 0.003, 0.003, 0.9, B.BRANCH, device=local)
    chain = RevJumpChain(data, model, seed=SEED, step=0.02, prior=IRRevJumpPrior(Tl=model.treelength())).init_()
    chain.mcmc_(5, trace=0)
    data.synchronize()
    p0 = chain.proposed
    gens = 60
    t0 = time.perf_counter()
    chain.mcmc_(gens, trace=0)
    data.synchronize()
    dt = time.perf_counter() - t0
    return {"metric": "mcmc! generations/sec", "value": gens / dt, "unit": "generations/s", "proposals_per_s": (chain.proposed - p0) / dt,
            "acceptance": chain.accepted / max(chain.proposed, 1), "logp": chain.state["logp"],
            "config": {"workload": "C2: 12monocots.nw + 12monocots-f01-1000.csv (%d patterns), fixed-dimension mcmc! under IRRevJumpPrior, 1 GPU" % len(data)}}


def algorithmic_fma(model, X):
    """Algorithmic fp64 FMAs of one full sweep over the patterns in X: per internal node with
    children counts M_1..M_k, S_i = M_1+…+M_i:  Σ_i [c_i internal ? (M_i+1)(M_i+2)/2 : 0]
    + Σ_{i>=2} [S_{i-1}(M_i+1) + (S_{i-1}+1)(M_i+1)] + Σ_i (S_i+1);  the root adds 3·x_root."""
    from beluga_b200.model import nonwgdchild
    nb = X.shape[1]

    def col(n):                       # inserted WGD nodes alias the counts of the species node below them
        return X[:, (n.i if n.i <= nb else nonwgdchild(n).i) - 1].astype(np.float64)

    tot = 0.0
    for n in model.nodes.values():
        if not n.c:
            continue
        S = None
        for i, c in enumerate(n.c):
            M = col(c)
            if c.c:
                tot += float(((M + 1) * (M + 2) / 2).sum())
            if i >= 1:
                tot += float((S * (M + 1) + (S + 1) * (M + 1)).sum())
            S = M if S is None else S + M
            tot += float((S + 1).sum())
    tot += 3.0 * float(X[:, 0].astype(np.float64).sum())
    return tot


def node_local_leg(args, local):
    """Config C3: 9dicots + 3 WGDs (23 nodes), 10^5 synthetic families; `logpdf!(model[i], data)` for
    every node i through the public API (θ update on the host, ε/W rebuild of the path, partial sweep, scalar back),
    then `rev!`; reports the mean over nodes and the time of `set!`/`rev!`."""
    import beluga_b200 as B
    from beluga_b200.model import build_model, update_
    from beluga_b200.sim import rand_profiles_device

    nw = open(os.path.join(ROOT, "tests", "data", "9dicots.nw")).readline().strip()
    m0, _, _ = build_model(nw, 1.0, 0.92, 0.66, B.BRANCH)
    for i, t, q in ((6, 0.01, 0.25), (14, 0.02, 0.5), (10, 0.05, 0.3)):
        m0.addwgd_(m0[i], t, q)
    names, counts = rand_profiles_device(m0, 100_000, seed=SEED, device=local, max_rootsum=60)
    model, data = B.DLWGD(nw, (names, counts), 1.0, 0.92, 0.66, B.BRANCH, device=local)
    for i, t, q in ((6, 0.01, 0.25), (14, 0.02, 0.5), (10, 0.05, 0.3)):
        model.addwgd_(model[i], t, q)
        B.extend_(data, i)
    B.logpdf_(model, data)
    B.set_(data)
    ids = sorted(i for i, n in model.nodes.items() if n.p is not None and "λ" in n.theta)
    per = {}
    reps = 5
    for i in ids:
        n = model[i]
        lam = n["λ"]
        for r in range(reps + 1):
            if r == 1:
                data.synchronize()
                t0 = time.perf_counter()
            update_(n, λ=lam * (1.0 + 1e-4 * (r + 1)), μ=n["μ"])
            B.logpdf_(n, data)
            update_(n, λ=lam, μ=n["μ"])
            B.rev_(data)
        per[i] = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(200):
        B.set_(data)
        B.rev_(data)
    cr = (time.perf_counter() - t0) / 400
    mean = float(np.mean(list(per.values())))
    return {"metric": "family-likelihood evals/sec (logpdf!(model[i]) node-local recompute)", "value": len(data) / mean, "unit": UNIT,
            "config": {"workload": "C3: 9dicots + 3 WGDs (%d nodes), 100000 synthetic families (%d patterns); update! + logpdf!(model[i]) + rev! for every node i, host in / host out" % (len(model), len(data))},
            "mean_ms": mean * 1e3, "min_ms": min(per.values()) * 1e3, "max_ms": max(per.values()) * 1e3,
            "commit_or_revert_us": cr * 1e6}


def gradient_leg(args, local):
    """`gradient(model, data)`: C1 (README example, 9dicots × 25 families, P = 35) and the C4 tree with 10^4 families."""
    import beluga_b200 as B
    from beluga_b200.model import build_model
    from beluga_b200.sim import rand_profiles_device
    out = {}
    nw = open(os.path.join(ROOT, "tests", "data", "9dicots.nw")).readline().strip()
    rows = [ln.strip().split(",") for ln in open(os.path.join(ROOT, "tests", "data", "9dicots-f01-25.csv"))]
    names, counts = rows[0], np.array(rows[1:], dtype=np.int64)
    model, data = B.DLWGD(nw, (names, counts), 1.0, 0.92, 0.66, B.BRANCH, device=local)
    B.gradient(model, data)
    t0 = time.perf_counter()
    for _ in range(3):
        g = B.gradient(model, data)
    out["C1"] = {"ms": (time.perf_counter() - t0) / 3 * 1e3, "P": int(len(g)), "patterns": len(data)}
    nw = open(os.path.join(ROOT, "tests", "data", "12monocots.nw")).readline().strip()
    m0, _, _ = build_model(nw, 0.004, 0.003, 0.9, B.BRANCH)
    names, counts = rand_profiles_device(m0, 10_000, seed=SEED, device=local, max_rootsum=60)
    model, data = B.DLWGD(nw, (names, counts), 0.004, 0.003, 0.9, B.BRANCH, device=local)
    B.gradient(model, data)
    t0 = time.perf_counter()
    g = B.gradient(model, data)
    out["C4-tree"] = {"ms": (time.perf_counter() - t0) * 1e3, "P": int(len(g)), "patterns": len(data)}
    return out


def ncu_traffic():
    """dram read+write bytes of the pruning launches of ONE sweep of this workload, summed over the launches, from the
    committed ncu capture (profiles/r02_sweep_traffic.json, which names the commit it was taken at); None if missing."""
    path = os.path.join(ROOT, "profiles", "r02_sweep_traffic.json")
    try:
        d = json.load(open(path))
        return {"bytes": d["dram_bytes_per_sweep"], "commit": d.get("commit"), "launches": d.get("launches")}
    except (OSError, ValueError, KeyError):
        return None


def workload_config(args, npat, nfam):
    return {"workload": "C5, reading 'root-sum <= %d' (m = %d, the reference's m = max of the extended profile = root count, src/model.jl:25,48): "
                        "%d-taxon coalescent tree + %d WGDs, %s families/GPU drawn from the model (rand(model, N), condition = rootclades), "
                        "pattern-compressed; full logpdf! sweep.  The other reading ('leaf count <= %d', m data-determined) is the c5_leaf50 block."
                        % (args.max_count, args.max_count + 1, args.taxa, args.wgds, args.families, args.max_count),
            "unit_note": "a 'family' of the metric is one pattern-compressed family (a unique profile row, what logpdf! evaluates); "
                         "weighted_families_per_s counts the multiplicities",
            "patterns_per_gpu": npat, "families_per_gpu": nfam, "tree_seed": SEED, "lambda": 1.0, "mu": 1.2, "eta": 0.8,
            "model": "Branch", "l2": "inputs larger than L2: a sweep writes and re-reads ~0.3 GB of partial-likelihood columns (126 MB L2)"}


def leaf50_leg(args, local):
    """C5 in the north star's primary reading: families are rejected only when a LEAF count exceeds 50; m is
    data-determined (root sums in the hundreds to thousands).  Families above the table-free kernel's count bound run on the
    general kernel (one warp per family, fp64-bound); the others on the pattern path."""
    import beluga_b200 as B
    from beluga_b200.model import build_model, nonwgdchild
    from beluga_b200.sim import add_random_wgds, rand_profiles_device, random_tree_newick
    nw = random_tree_newick(args.taxa, seed=SEED)
    model, t, l = build_model(nw, 1.0, 1.2, 0.8, B.BRANCH)
    ext = add_random_wgds(model, args.wgds, seed=SEED)
    names, counts = rand_profiles_device(model, args.leaf50_families, seed=SEED, device=local, max_leaf=args.max_count)
    X, w, m = B.profile(t, l, (names, counts))
    fma = algorithmic_fma(model, X)
    data = B.PArray(X, w, device=local)
    for i in ext:
        B.extend_(data, i)
    B.set_(data)
    B.logpdf_(model, data)
    data.set_timing(True)
    reps = 3
    data.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ll = B.logpdf_(model, data)
    dt = (time.perf_counter() - t0) / reps
    pm, rm = data.last_timing()
    _, _, nfast, nheavy = data.kernel_runs()
    alg_bytes, prune_bytes, launches = data.last_algorithmic()
    peak = data.fp64_peak()
    root = X[:, 0]
    return {"workload": "C5, reading 'leaf count <= %d' (m = %d data-determined): %d families drawn from the model, %d patterns" % (args.max_count, m, args.leaf50_families, X.shape[0]),
            "value": X.shape[0] / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "prune_ms": pm, "logpdf": ll,
            "root_count_quantiles": {"min": int(root.min()), "median": float(np.median(root)), "q99": float(np.quantile(root, 0.99)), "max": int(root.max())},
            "patterns_on_pattern_path": int(nfast), "patterns_on_general_kernel": int(nheavy),
            "roofline_fp64": {"bound": "fp64", "achieved": 2.0 * fma / (pm / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                              "frac": 2.0 * fma / (pm / 1e3) / 1e12 / peak, "algorithmic_fma_per_step": fma},
            "roofline_hbm": {"achieved": prune_bytes / 1e9 / (pm / 1e3), "unit": "GB/s", "algorithmic_bytes": prune_bytes}}


def strong_scaling_leg(args, rank, world, local):
    """ONE draw of `--families` families (rank 0's seed), compressed globally, sharded contiguously over the ranks by the
    product's own path (DLWGD_(distributed=True) semantics: profile() then shard_range), evaluated with the library's own
    communicator.  Reports ms per full sweep, resident and end to end."""
    import torch
    import torch.distributed as dist
    import beluga_b200 as B
    nw, model, ext, t, l, names, counts = build_workload(args, 0, device="cuda:%d" % local)
    X, w, m = B.profile(t, l, (names, counts))
    lo, hi = B.shard_range(X.shape[0], rank, world)
    Xs = np.ascontiguousarray(X[lo:hi], dtype=np.int32)
    data = B.PArray(Xs, w[lo:hi], device=local, distributed=True)
    for i in ext:
        B.extend_(data, i)
    B.set_(data)
    total = B.logpdf_(model, data)
    stream = data.torch_stream
    import ctypes as C
    out = torch.zeros(1, dtype=torch.float64, device="cuda:%d" % local)
    ptr = C.c_void_p(out.data_ptr())
    with torch.cuda.stream(stream):
        for _ in range(3):
            data._ck(data.L.blg_logpdf_full_async(data.h, ptr))
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            data._ck(data.L.blg_logpdf_full_async(data.h, ptr))
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    rates = model.getrates()
    rng = np.random.default_rng(SEED)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    esteps = max(3, args.steps // 2)
    for _ in range(esteps):
        model.setrates_(rates * np.exp(rng.normal(0, 1e-3, size=rates.shape)))
        B.logpdf_(model, data)
    dist.barrier(); torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / esteps
    tt = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda:%d" % local)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms = [float(x) for x in tt.tolist()]
    return {"workload": "ONE %d-family draw, %d patterns after global compression, sharded contiguously over %d GPUs by the library's path" % (args.families, X.shape[0], world),
            "scaling": "strong", "patterns_total": int(X.shape[0]), "patterns_this_rank": int(hi - lo), "ms_per_step": ms, "value": X.shape[0] / (ms / 1e3), "unit": UNIT,
            "e2e_ms_per_step": e2e_ms, "e2e_value": X.shape[0] / (e2e_ms / 1e3), "logpdf": total}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU algorithm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()

    import beluga_b200 as B

    nw, model, ext, t, l, names, counts = build_workload(args, rank, device="cuda:%d" % local)
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_i0 = time.perf_counter()
    X, w, m = B.profile(t, l, (names, counts))
    t_i1 = time.perf_counter()
    npat, nfam = X.shape[0], int(w.sum())
    alg_fma = algorithmic_fma(model, X)
    t_i2 = time.perf_counter()
    data = B.PArray(X, w, device=local, distributed=True)
    t_i3 = time.perf_counter()
    for i in ext:                 # extend!(data, i) for every inserted WGD (src/profile.jl:108), then set!(data)
        B.extend_(data, i)
    B.set_(data)
    B.logpdf_(model, data)        # first call: ε/W tables, node patterns, links, slabs
    t_i4 = time.perf_counter()
    ingest = {"what": "host wall clock of getting the count table (%d x %d, %s) ready for the first likelihood on rank 0: profile() = site patterns "
                      "on the device + extended counts; PArray = blg_set_profiles_columns; first logpdf! = extend!/set!, ε/W tables, node "
                      "patterns, links, slabs, one sweep" % (counts.shape[0], counts.shape[1], counts.dtype),
              "profile_s": t_i1 - t_i0, "set_profiles_s": t_i3 - t_i2, "first_logpdf_s": t_i4 - t_i3}
    # the CPU baseline gets a uniform sample of THIS pattern set (same pattern mix as the CUDA arm)
    leaf_ids = sorted(l)
    sel = np.random.default_rng(SEED).permutation(X.shape[0])[:20000]
    cpu_sample = (nw, [l[i] for i in leaf_ids], X[sel][:, [i - 1 for i in leaf_ids]].astype(np.int64),
                  "uniform sample of the %d patterns of rank 0's compressed %d-family set" % (X.shape[0], args.families), True)
    del X, counts

    stream = data.torch_stream
    torch.cuda.set_stream(stream)          # everything below (kernels, all-reduce, events) on the handle's stream
    out = torch.zeros(1, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernel-only: inputs resident, parameters fixed ------------------------------------------------
    data.sync(model)
    import ctypes as C
    ptr = C.c_void_p(out.data_ptr())

    def step_resident():
        data._ck(data.L.blg_logpdf_full_async(data.h, ptr))      # (sweep, root, and the library's own all-reduce of the scalar)

    sampler = ClockSampler(local, enabled=(rank == 0))     # samples clocks from the warm-up through both timed regions
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    total_ll = float(out.item())
    graph_captures, graph_replays = data.graph_stats()
    # per-launch-group durations: the same step, the same number of times, with the library's own events around the pruning
    # launches and the root kernel (event-record nodes inside the replayed graph: they cut it into three launches, which is
    # why this loop is not the one `value` is taken from); mean over the steps
    data.set_timing(True)
    for _ in range(3):
        step_resident()
    pms, rms = [], []
    for _ in range(args.steps):
        step_resident()
        a_, b_ = data.last_timing()
        pms.append(a_); rms.append(b_)
    pm, rm = float(np.mean(pms)), float(np.mean(rms))
    alg_bytes, prune_bytes, launches = data.last_algorithmic()
    moved_bytes, pattern_cols, family_cols = data.pattern_stats()
    data.set_timing(False)

    # ---- end to end through the public API: host θ in, host scalar out, every step -------------------------
    rates = model.getrates()
    rng = np.random.default_rng(SEED)
    for _ in range(max(1, args.warmup // 2)):
        model.setrates_(rates * np.exp(rng.normal(0, 1e-3, size=rates.shape)))
        B.logpdf_(model, data)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record(stream)
    esteps = max(3, args.steps // 2)
    for _ in range(esteps):
        model.setrates_(rates * np.exp(rng.normal(0, 1e-3, size=rates.shape)))   # host-side parameter update
        ll = B.logpdf_(model, data)                                             # H2D θ, ε/W build, sweep, all-reduce, D2H scalar
    e3.record(stream)
    barrier()
    e2e_ms = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3) / esteps
    sampler.stop_.set()
    sampler.join(2)
    ncols = max(model.nodes)
    h2d = 4 * ncols * 8 + ncols * 176 + ncols * 4
    model.setrates_(rates)

    # ---- max over ranks ----------------------------------------------------------------------------------
    tt = torch.tensor([ms, e2e_ms, pm, rm], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(npat), float(nfam), alg_bytes], dtype=torch.float64, device="cuda")
    # (roofline numbers below are rank 0's own kernel; counts are summed over ranks)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_ms, pm, rm = [float(x) for x in tt.tolist()]
    tot_pat, tot_fam, tot_bytes = [float(x) for x in cnt.tolist()]

    # ---- legs every rank takes part in: the sharded 10^6-family problem (strong scaling), rjMCMC on sharded data ----------
    strong = None
    mcmc_n = None
    if world > 1:
        torch.cuda.set_stream(torch.cuda.default_stream())
        try:
            strong = strong_scaling_leg(args, rank, world, local)
        except Exception as e:
            strong = {"error": repr(e)}
        if not args.no_mcmc:
            try:
                mcmc_n = mcmc_leg(args, local, distributed=True)
            except Exception as e:
                mcmc_n = {"metric": "rjMCMC iters/sec", "value": None, "error": repr(e)}
        dist.barrier()

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel: sweep_kernel (one launch per likelihood call; the other launch is root_kernel).  Its algorithmic
        # bytes are the sweep's minus the root/reduction share; its duration is the event-bracketed pruning phase.
        achieved = (prune_bytes / 1e9) / (pm / 1e3)
        line = {
            "metric": METRIC, "value": tot_pat / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, npat, nfam),
            "weighted_families_per_s": tot_fam / (ms / 1e3),
            "logpdf": total_ll,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (ncu_traffic() or {}).get("bytes"), "traffic_source": ncu_traffic(),
                         "kernel": "the pruning launches of one sweep (ptile_kernel per tree level, %d launches/step, event-bracketed as one group)" % (launches - 1),
                         "algorithmic_bytes_per_launch_group": prune_bytes, "algorithmic_bytes_per_step": alg_bytes, "phase_ms": {"prune": pm, "root+reduce": rm},
                         "phase_ms_note": "mean over %d further steps of the same resident call with the library's events around the pruning launches "
                                          "(they split the replayed CUDA graph, so `ms_per_step` comes from the loop without them)" % args.steps,
                         "peak_source": peak_src,
                         "note": "achieved = SURVEY 8(d)'s algorithmic bytes (one column per FAMILY per node, written once and read once) / time. "
                                 "The kernels evaluate one column per distinct NODE PATTERN (families deduplicated per node by the counts below it), so the "
                                 "bytes actually moved (pattern_path.bytes_moved_model, and `traffic` = measured DRAM bytes) are far below the algorithmic figure"},
            "pattern_path": {"node_steps_x_families": family_cols, "node_steps_x_patterns": pattern_cols,
                             "dedup_factor": (family_cols / pattern_cols) if pattern_cols else None, "bytes_moved_model": moved_bytes},
            "e2e": {"value": tot_pat / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms, "what": "setrates!(model, X) on host + logpdf!(model, profile): θ upload, on-device ε/W build, full sweep, scalar back"},
            "gpu_launches": int(launches * args.steps),
            "cuda_graph": {"captures": int(graph_captures), "replays": int(graph_replays),
                           "note": "kernel launches of a resident step are replayed from one captured graph (blg_graph_stats); gpu_launches counts the kernels inside"},
            "clocks": sampler.summary(),
        }
        try:                                 # the same step against the fp64 roofline (evidence of the regime)
            fp64_peak = data.fp64_peak()
            tf = 2.0 * alg_fma / (pm / 1e3) / 1e12
            line["roofline_fp64"] = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                                     "algorithmic_fma_per_step": alg_fma, "flop_per_byte": 2.0 * alg_fma / prune_bytes,
                                     "peak_source": "measured on this GPU (blg_fp64_peak: 8 DFMA chains/thread)"}
        except Exception as e:
            line["roofline_fp64"] = {"error": repr(e)}
        c5_grad = None
        if not args.no_mcmc and world == 1:
            try:                             # gradient(model, data) on the headline workload itself (reverse-mode sweep)
                data.gradient(model)
                t0 = time.perf_counter()
                gg, _ = data.gradient(model)
                c5_grad = {"ms": (time.perf_counter() - t0) * 1e3, "P": int(len(gg)), "patterns": int(npat)}
            except Exception as e:
                c5_grad = {"error": repr(e)}
        if not args.no_mcmc:
            for key, leg in (("node_local", node_local_leg), ("gradient", gradient_leg), ("c2_mcmc", c2_leg)):
                try:
                    line[key] = leg(args, local)
                except Exception as e:
                    line[key] = {"error": repr(e)}
            if c5_grad is not None and isinstance(line.get("gradient"), dict):
                line["gradient"]["C5"] = c5_grad
        line["ingest"] = ingest
        if strong is not None:
            line["strong_scaling"] = strong
        if world == 1 and args.leaf50_families > 0:
            try:
                line["c5_leaf50"] = leaf50_leg(args, local)
            except Exception as e:
                line["c5_leaf50"] = {"error": repr(e)}
        if mcmc_n is not None:
            line["mcmc"] = mcmc_n
        elif not args.no_mcmc:
            try:
                del data
                torch.cuda.empty_cache()
                line["mcmc"] = mcmc_leg(args, local)
            except Exception as e:
                line["mcmc"] = {"metric": "rjMCMC iters/sec", "value": None, "error": repr(e)}
        if not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args, None, args.cpu_seconds, sample=cpu_sample).items()
                                        if k in ("value", "unit", "cores", "kind", "sample", "same_pattern_mix")}
            except Exception as e:  # the baseline must not take the measurement down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
