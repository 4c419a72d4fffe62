"""Synthetic inputs for tests and benchmarks.

``rand_profiles`` is a COUNT-LEVEL sampler of the DLWGD model, equal in distribution to what
``rand(model, N)`` (src/sim.jl:17-49) is meant to draw: root lineages ~ Geometric(η)+1
(sim.jl:78), each lineage evolving along each branch under the linear birth–death process with the
branch's (λ, μ) (sim.jl:124-145, here through its closed-form transition law: extinct w.p. ϕ(t),
otherwise 1 + Geometric(1-ψ(t)) descendants, utils.jl:8-9), each lineage doubling w.p. q at a WGD
node (sim.jl:112-115).  One deliberate difference: the reference's gene-tree simulator skips the
branch segment between a WGD and the next speciation (sim.jl:94-96 jumps over the wgdafter node
without simulating the child's remaining branch); this sampler simulates that segment, i.e. it
draws from the model whose likelihood the hot path evaluates.  Benchmark-input generation is not on
the hot path."""
import numpy as np

from .model import isawgd, isleaf, iswgd, iswgdafter, iswgt, nonwgdchild, nonwgdparent, prewalk


def random_tree_newick(ntaxa, seed=20261019, height=1.0, prefix="t"):
    """Kingman-coalescent topology and times, rescaled to the given root height (ultrametric)."""
    rng = np.random.default_rng(seed)
    nodes = [("%s%03d" % (prefix, i), 0.0) for i in range(ntaxa)]   # (newick string without length, height)
    k = ntaxa
    tnow = 0.0
    events = []
    live = list(range(ntaxa))
    sub = {i: (nodes[i][0], 0.0) for i in live}
    nxt = ntaxa
    while k > 1:
        tnow += rng.exponential(1.0 / (k * (k - 1) / 2.0))
        a, b = rng.choice(len(live), size=2, replace=False)
        ia, ib = live[a], live[b]
        events.append((ia, ib, tnow))
        sub[nxt] = (None, tnow, ia, ib)
        live = [x for j, x in enumerate(live) if j not in (a, b)] + [nxt]
        nxt += 1
        k -= 1
    scale = height / tnow

    def render(i, parent_h):
        rec = sub[i]
        if rec[0] is not None:
            return "%s:%.10f" % (rec[0], (parent_h - 0.0) * scale)
        _, h, a, b = rec
        inner = "(%s,%s)" % (render(a, h), render(b, h))
        if parent_h is None:
            return inner + ";"
        return "%s:%.10f" % (inner, (parent_h - h) * scale)

    return render(nxt - 1, None)


def _phi_psi(t, lam, mu):
    if abs(lam - mu) <= 1.4901161193847656e-8 * max(abs(lam), abs(mu)):
        v = lam * t / (1.0 + lam * t)
        return v, v
    e = np.exp(t * (lam - mu))
    phi = mu * (e - 1.0) / (lam * e - mu)
    return phi, (lam / mu) * phi


def _branch_rates(model, parent, child):
    """getλμ(nonwgdparent(node), nonwgdchild(c)) as src/sim.jl:108 calls it (src/node.jl:220-230)."""
    a, b = nonwgdparent(parent), nonwgdchild(child)
    if model.nt == "Node":
        return 0.5 * (a["λ"] + b["λ"]), 0.5 * (a["μ"] + b["μ"])
    return b["λ"], b["μ"]


def _evolve(rng, n, t, lam, mu):
    """n lineages through a branch of length t under the linear BDP: vector in, vector out."""
    if t <= 0.0:
        return n
    phi, psi = _phi_psi(t, lam, mu)
    surv = rng.binomial(n, 1.0 - phi)
    extra = np.zeros_like(n)
    pos = surv > 0
    if psi > 0.0 and pos.any():
        extra[pos] = rng.negative_binomial(surv[pos], 1.0 - psi)
    return surv + extra


def rand_counts(model, N, rng):
    """N unconditioned draws; returns (leaf_ids, counts[N × nleaves])."""
    root = model.nodes[1]
    at = {1: rng.geometric(root["η"], size=N).astype(np.int64)}
    for n in prewalk(root):
        x = at[n.i]
        for c in n.c:
            if iswgt(n):
                raise NotImplementedError("sampling through WGT nodes is not supported")
            y = x
            if iswgd(n):
                y = x + rng.binomial(x, n["q"])          # retention of the duplicate w.p. q
            lam, mu = _branch_rates(model, n, c)
            at[c.i] = _evolve(rng, y, c["t"], lam, mu)
    leaf_ids = sorted(i for i, nd in model.nodes.items() if isleaf(nd))
    return leaf_ids, np.stack([at[i] for i in leaf_ids], axis=1)


def rootclades(model):
    """Beluga.rootclades(d) (src/sim.jl:51): leaf ids below each child of the root."""
    out = []
    for c in model.nodes[1].c:
        out.append(sorted(n.i for n in prewalk(c) if isleaf(n)))
    return out


def rand_profiles(model, N, seed=20261019, condition="rootclades", max_leaf=None, max_rootsum=None):
    """rand(model, N, condition=…): rejection-sample N profiles.  ``condition``: "rootclades" (at least one
    gene in each clade stemming from the root), "nonextinct" (the reference's default) or None.
    max_leaf / max_rootsum additionally reject families with a leaf count / total count above the bound.
    -> (names sorted alphabetically like DataFrame(sort(profile)), counts[N × ntaxa])."""
    rng = np.random.default_rng(seed)
    out = []
    have = 0
    leaf_ids = None
    clades = rootclades(model) if condition == "rootclades" else None
    while have < N:
        leaf_ids, c = rand_counts(model, max(1024, int((N - have) * 1.5)), rng)
        keep = np.ones(c.shape[0], dtype=bool)
        if condition == "rootclades":
            pos = {i: k for k, i in enumerate(leaf_ids)}
            for cl in clades:
                keep &= c[:, [pos[i] for i in cl]].sum(axis=1) > 0
        elif condition == "nonextinct":
            keep &= c.sum(axis=1) > 0
        if max_leaf is not None:
            keep &= c.max(axis=1) <= max_leaf
        if max_rootsum is not None:
            keep &= c.sum(axis=1) <= max_rootsum
        c = c[keep]
        out.append(c)
        have += c.shape[0]
    counts = np.concatenate(out, axis=0)[:N]
    names = [model.leaves[i] for i in leaf_ids]
    order = np.argsort(names)
    return [names[k] for k in order], counts[:, order]


# ---- benchmark-scale input generation on the GPU (same law as rand_counts, torch RNG) ------------------
def rand_profiles_torch(model, N, seed=20261019, device="cuda", condition="rootclades", max_leaf=None,
                        max_rootsum=None, chunk=1 << 21):
    """Rejection sampler equal in distribution to ``rand_profiles`` that draws on the device with
    torch's generators (input generation only; not part of the likelihood path).  Returns
    (names, counts[N × ntaxa] int16 ndarray)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    root = model.nodes[1]
    order = prewalk(root)
    leaf_ids = sorted(i for i, nd in model.nodes.items() if isleaf(nd))
    pos = {i: k for k, i in enumerate(leaf_ids)}
    clades = rootclades(model) if condition == "rootclades" else None
    out, have = [], 0
    while have < N:
        at = {1: torch.empty(chunk, dtype=torch.float64, device=device).geometric_(float(root["η"]), generator=gen)}
        leaves = torch.empty((chunk, len(leaf_ids)), dtype=torch.int16, device=device)
        for n in order:
            x = at.pop(n.i)
            if isleaf(n):
                leaves[:, pos[n.i]] = x.to(torch.int16)
                continue
            for c in n.c:
                if iswgt(n):
                    raise NotImplementedError("sampling through WGT nodes is not supported")
                y = x
                if iswgd(n):
                    y = x + torch.binomial(x, torch.full_like(x, float(n["q"])), generator=gen)
                lam, mu = _branch_rates(model, n, c)
                t = c["t"]
                if t > 0.0:
                    phi, psi = _phi_psi(t, lam, mu)
                    surv = torch.binomial(y, torch.full_like(y, 1.0 - float(phi)), generator=gen)
                    if psi > 0.0:
                        g = torch._standard_gamma(surv.clamp_min(1e-30), generator=gen) * (float(psi) / (1.0 - float(psi)))
                        y = surv + torch.poisson(g, generator=gen)
                    else:
                        y = surv
                at[c.i] = y
        keep = torch.ones(chunk, dtype=torch.bool, device=device)
        wide = leaves.to(torch.int32)
        if condition == "rootclades":
            for cl in clades:
                keep &= wide[:, [pos[i] for i in cl]].sum(dim=1) > 0
        elif condition == "nonextinct":
            keep &= wide.sum(dim=1) > 0
        if max_leaf is not None:
            keep &= wide.max(dim=1).values <= max_leaf
        if max_rootsum is not None:
            keep &= wide.sum(dim=1) <= max_rootsum
        sel = leaves[keep].cpu().numpy()
        out.append(sel)
        have += sel.shape[0]
    counts = np.concatenate(out, axis=0)[:N]
    names = [model.leaves[i] for i in leaf_ids]
    order_ = np.argsort(names)
    return [names[k] for k in order_], np.ascontiguousarray(counts[:, order_])


def rand_profiles_device(model, N, seed=20261019, device=0, condition="rootclades", max_leaf=None, max_rootsum=None):
    """rand(model, N, condition=…) drawn by the library's own sampler kernel (csrc/sampler.cuh: one thread and one Philox
    stream per family, rejection with ordered compaction): equal in distribution to ``rand_profiles``, reproducible
    for a given seed, WGT nodes supported.  Returns (names sorted alphabetically, counts[N × ntaxa] uint16)."""
    import ctypes as C
    from . import _lib
    L = _lib.lib()
    h = C.c_void_p()
    if L.blg_create(int(device), None, C.byref(h)) != 0:
        raise _lib.BelugaError(L.blg_last_error(None).decode())
    try:
        def ck(rc):
            if rc != 0:
                raise _lib.BelugaError(L.blg_last_error(h).decode())
        n, parent, kind, cpos = model.flat_topology()
        ck(L.blg_set_topology(h, n, parent.ctypes.data_as(C.POINTER(C.c_int32)), kind.ctypes.data_as(C.POINTER(C.c_uint8)),
                              cpos.ctypes.data_as(C.POINTER(C.c_int32))))
        lam, mu, q, t, eta = model.flat_params()
        dp = C.POINTER(C.c_double)
        ck(L.blg_set_params(h, 0 if model.nt == "Node" else 1, lam.ctypes.data_as(dp), mu.ctypes.data_as(dp), q.ctypes.data_as(dp),
                            t.ctypes.data_as(dp), float(eta)))
        nleaf = sum(1 for nd in model.nodes.values() if isleaf(nd))
        leaf_ids = np.zeros(n, dtype=np.int32)
        nl = C.c_int()
        out = np.zeros((int(N), nleaf), dtype=np.uint16)
        draws = C.c_longlong()
        cond = {None: 0, "nonextinct": 1, "rootclades": 2}[condition]
        ck(L.blg_rand_profiles(h, int(N), C.c_ulonglong(int(seed)), cond, -1 if max_leaf is None else int(max_leaf),
                               -1 if max_rootsum is None else int(max_rootsum), leaf_ids.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(nl),
                               out.ctypes.data_as(C.POINTER(C.c_uint16)), C.byref(draws)))
        assert nl.value == nleaf
    finally:
        L.blg_destroy(h)
    names = [model.leaves[int(i)] for i in leaf_ids[:nleaf]]
    order = np.argsort(names)
    rand_profiles_device.last_draws = draws.value
    return [names[k] for k in order], np.ascontiguousarray(out[:, order])


def add_random_wgds(model, k, seed=20261019, qa=1.0, qb=3.0):
    """k WGDs placed with the reference's randpos semantics (src/model.jl:197-206: branch ∝ length,
    uniform position along it), retention q ~ Beta(qa, qb).  Returns the ids extend! must be
    called with, in insertion order."""
    rng = np.random.default_rng(seed)
    ext = []
    for _ in range(k):
        ids = sorted(model.nodes)
        w = np.array([model.nodes[i]["t"] for i in ids], dtype=np.float64)
        i = int(rng.choice(ids, p=w / w.sum()))
        n = model.nodes[i]
        t = rng.random() * n["t"]
        if not (n["t"] - t > 0.0):
            t = 0.5 * n["t"]
        model.addwgd_(n, t, float(rng.beta(qa, qb)))
        ext.append(i)
    return ext
