"""Device-resident replacement of ``Profile``/``PArray`` (src/profile.jl).

The reference keeps, per family, the extended counts ``x``/``xp`` and the log-space DP matrices
``L``/``Lp`` in a DArray of heap objects and evaluates them with ``mapreduce``.  Here the same state
lives in one CUDA library handle (include/beluga_b200.h): counts and partial likelihoods are
node-major, family-contiguous device arrays, ``set!``/``rev!`` are pointer edits, and the likelihood
entries launch the sm_100a kernels.  Method-for-method:

    logpdf!(d, p), logpdf!(n, p)   profile.jl:56-60   -> logpdf_(d_or_n, p)
    logpdfroot(n, p)               profile.jl:62-63   -> logpdfroot(n, p)
    gradient, gradient_cr          profile.jl:71-75   -> gradient(d, p), gradient_cr(d, p)
    set!, rev!                     profile.jl:79-80   -> set_(p), rev_(p)
    extend!, shrink!               profile.jl:108-109 -> extend_(p, i), shrink_(p, i)
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .model import BRANCH, NODE, DLWGD, ModelNode, build_model, postwalk
from .parallel import allreduce_sum_, dist_info, shard_range

_MK = {NODE: 0, BRANCH: 1}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PArray:
    """PArray{Float64}: the phylogenetic profile matrix of one process / one GPU.

    X: (npatterns, ncols) extended counts (``permutedims(M)`` of model.jl:48, one row per site
    pattern), weights: pattern multiplicities.  ``X is None`` is the mock profile ``PArray()``
    (profile.jl:28,31) and needs no GPU.  With ``distributed=True`` the patterns passed in are this
    rank's shard; the library joins an NCCL communicator (the 128-byte id travels over
    torch.distributed, nothing else does) and every likelihood / gradient entry returns the sum over
    the ranks — the all-reduce is issued by the library on its own stream."""

    def __init__(self, X=None, weights=None, device=0, distributed=False):
        self.h = None
        self.L = None
        self.distributed = bool(distributed)
        self._model = None
        self._topo_seen = self._param_seen = self._seq_seen = -1
        self._torch_out = None
        if X is None:
            self.N = 0
            self.ncols0 = 0
            return
        X = np.asarray(X)
        # a column-major matrix (what profile() builds: one contiguous run per node) goes to the library as it is
        self._colmajor = X.ndim == 2 and X.dtype == np.int32 and X.flags.f_contiguous and not X.flags.c_contiguous
        if not self._colmajor:
            X = np.ascontiguousarray(X, dtype=np.int32)
        weights = np.ascontiguousarray(weights, dtype=np.int64)
        assert X.ndim == 2 and weights.shape == (X.shape[0],)
        self.N = X.shape[0]
        self.ncols0 = X.shape[1]
        self.L = _lib.lib()
        stream = None
        self.torch_stream = None
        if self.distributed:
            # a dedicated torch stream: the library's kernels, the NCCL all-reduce behind them and
            # the caller's CUDA events all go on it
            import torch
            torch.cuda.set_device(device)
            self.torch_stream = torch.cuda.Stream(device=device)
            stream = C.c_void_p(self.torch_stream.cuda_stream)
        h = C.c_void_p()
        rc = self.L.blg_create(int(device), stream, C.byref(h))
        if rc != 0:
            raise _lib.BelugaError(self.L.blg_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self.world = 1
        if self.distributed:
            self._join_communicator()
        entry = self.L.blg_set_profiles_columns if self._colmajor else self.L.blg_set_profiles
        self._ck(entry(self.h, X.ctypes.data_as(C.POINTER(C.c_int32)), weights.ctypes.data_as(C.POINTER(C.c_int64)),
                       X.shape[1], X.shape[0]))

    def _join_communicator(self):
        """blg_comm_unique_id on rank 0, broadcast of the 128 bytes over torch.distributed, blg_comm_init everywhere."""
        import torch
        import torch.distributed as dist
        rank, world = dist_info()
        if world <= 1:
            return
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0 and self.L.blg_comm_unique_id(uid.ctypes.data_as(C.POINTER(C.c_uint8))) != 0:
            raise _lib.BelugaError(self.L.blg_last_error(None).decode())
        backend = dist.get_backend()
        t = torch.from_numpy(uid)
        if backend == "nccl":
            t = t.to("cuda:%d" % self.device)
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy().copy()
        self._ck(self.L.blg_comm_init(self.h, world, rank, uid.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.world = world

    def __del__(self):
        try:
            if self.h is not None:
                self.L.blg_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return self.N

    def __repr__(self):
        return "PArray{Float64}(%d)" % self.N

    def _ck(self, rc):
        if rc != 0:
            raise _lib.BelugaError(self.L.blg_last_error(self.h).decode())

    # ---- model → device mirroring -----------------------------------------------------------------
    def sync(self, model):
        """Replay on the device what the reference did eagerly on the host since the last call:
        topology edits, θ changes and the set!(n) requests recorded by update!/set!."""
        if self.h is None:
            return
        full = False
        if self._model is not model or self._topo_seen != model._topo_version:
            n, parent, kind, cpos = model.flat_topology()
            self._ck(self.L.blg_set_topology(self.h, n, parent.ctypes.data_as(C.POINTER(C.c_int32)),
                                             kind.ctypes.data_as(C.POINTER(C.c_uint8)),
                                             cpos.ctypes.data_as(C.POINTER(C.c_int32))))
            full = True
            self._param_seen = -1
        if self._param_seen != model._param_version or full:
            lam, mu, q, t, eta = model.flat_params()
            self._ck(self.L.blg_set_params(self.h, _MK[model.nt], _dp(lam), _dp(mu), _dp(q), _dp(t), eta))
        if self._model is not model or self._seq_seen < model._log_start:
            full = True
        ids = []
        for seq, req in model._log:
            if seq <= self._seq_seen:
                continue
            if req is None:
                full = True
            else:
                ids.extend(req)
        if full:
            self._ck(self.L.blg_rebuild(self.h, None, 0))
        elif ids:
            arr = np.asarray(ids, dtype=np.int32)
            self._ck(self.L.blg_rebuild(self.h, arr.ctypes.data_as(C.POINTER(C.c_int32)), len(arr)))
        self._model = model
        self._topo_seen = model._topo_version
        self._param_seen = model._param_version
        self._seq_seen = model._seq

    # ---- likelihood -----------------------------------------------------------------------------------
    def _run(self, mode, node):
        # (with a communicator the library has already summed over the ranks)
        out = C.c_double()
        if mode == 0:
            self._ck(self.L.blg_logpdf_full(self.h, C.byref(out)))
        elif mode == 1:
            self._ck(self.L.blg_logpdf_from(self.h, node, C.byref(out)))
        else:
            self._ck(self.L.blg_logpdf_root(self.h, C.byref(out)))
        return out.value

    def logpdf_(self, x):
        if self.N == 0 and not self.distributed:
            return 0.0                                                   # profile.jl:56
        if isinstance(x, DLWGD):
            self.sync(x)
            return self._run(0, 1)
        self.sync(x.model)
        return self._run(1, x.i)

    def logpdfroot(self, n):
        if self.N == 0 and not self.distributed:
            return 0.0
        self.sync(n.model)
        return self._run(2, 1)

    # ---- gradient (src/profile.jl:71-75) -----------------------------------------------------------------------
    def gradient(self, model, cr=False, literal=False):
        """-> (g, logpdf).  g is laid out like asvector(model) (src/model.jl:541-546): [λ…, μ…, q in getwgds
        order, η]; ``cr=True`` gives gradient_cr's [∂λ, ∂μ] with the rates tied across nodes.

        With two or more WGDs the reference's own ``gradient`` differentiates a model whose q's are permuted:
        asvector lists them by descending wgd id, model(θ) (src/model.jl:320-322) reads them by ascending id.  The
        default here is the gradient AT the model's parameters; ``literal=True`` reproduces the reference's
        evaluation (q's reversed among the WGDs, entry j of the q block = ∂/∂q of the j-th WGD by ascending id)."""
        n = model.ne() + 1
        k = model.nwgd()
        if literal and not cr and k > 1:
            wg = model.getwgds()                                    # descending ids
            qs = [model[w]["q"] for w in wg]
            for j, w in enumerate(wg[::-1]):                         # the j-th ascending WGD gets the j-th descending q
                model[w]["q"] = qs[j]
            model.set_()
            try:
                g, llv = self.gradient(model, cr=False, literal=False)
            finally:
                for w, q in zip(wg, qs):
                    model[w]["q"] = q
                model.set_()
            g[2 * n:2 * n + k] = g[2 * n:2 * n + k][::-1].copy()   # back to the library's ascending order = θ's positions
            return g, llv
        P = 2 if cr else 2 * n + k + 1
        g = np.zeros(P)
        if self.N == 0 and not self.distributed:
            return g, 0.0
        self.sync(model)
        ll = C.c_double()
        if cr:
            self._ck(self.L.blg_gradient_cr(self.h, _dp(g), C.byref(ll)))
        else:
            self._ck(self.L.blg_gradient(self.h, _dp(g), P, C.byref(ll)))
        llv = ll.value               # (summed over the ranks by the library when it owns a communicator)
        if not cr and k > 1:
            # the library orders ∂/∂q by ascending wgd id; asvector lists q in getwgds order (descending id)
            g[2 * n:2 * n + k] = g[2 * n:2 * n + k][::-1].copy()
        return g, llv

    # ---- accept / reject / reversible jump ---------------------------------------------------------------
    def set_(self):
        if self.h is not None:
            self._ck(self.L.blg_commit(self.h))

    def rev_(self):
        if self.h is not None:
            self._ck(self.L.blg_revert(self.h))

    def extend_(self, i):
        if self.h is not None:
            self._ck(self.L.blg_extend(self.h, int(i)))

    def shrink_(self, i):
        if self.h is not None:
            self._ck(self.L.blg_shrink(self.h, int(i)))

    # ---- inspection ---------------------------------------------------------------------------------------
    def column(self, pattern, node, proposal=True):
        cap = 4200
        out = np.full(cap, np.nan)
        rows = C.c_int()
        self._ck(self.L.blg_get_column(self.h, int(pattern), int(node), 1 if proposal else 0, _dp(out), cap, C.byref(rows)))
        return out[: rows.value]

    def family_logpdf(self):
        out = np.zeros(self.N)
        self._ck(self.L.blg_get_family_logpdf(self.h, _dp(out), self.N))
        return out

    def eps(self):
        n = self.L.blg_ncols(self.h, 1) if self._model is None else max(self._model.nodes)
        e0, e1 = np.zeros(n), np.zeros(n)
        self._ck(self.L.blg_get_eps(self.h, _dp(e0), _dp(e1), n))
        return e0, e1

    def W(self, node):
        dim = C.c_int()
        cap = 1024 * 1024
        out = np.zeros(cap)
        self._ck(self.L.blg_get_W(self.h, int(node), _dp(out), cap, C.byref(dim)))
        d = dim.value
        return out[: d * d].reshape(d, d).copy()

    def ncols(self, proposal=True):
        return self.L.blg_ncols(self.h, 1 if proposal else 0)

    def last_algorithmic(self):
        b, bp, l = C.c_double(), C.c_double(), C.c_int()
        self.L.blg_last_algorithmic(self.h, C.byref(b), C.byref(bp), C.byref(l))
        return b.value, bp.value, l.value

    def kernel_runs(self):
        """-> (sweep_runs, general_runs, nfast, nheavy): cumulative launches of each pruning kernel that did the work,
        and the split of the patterns between the fast and the heavy family set (diagnostic)."""
        a, b, nf, nh = C.c_uint(), C.c_uint(), C.c_int(), C.c_int()
        self._ck(self.L.blg_kernel_runs(self.h, C.byref(a), C.byref(b), C.byref(nf), C.byref(nh)))
        return a.value, b.value, nf.value, nh.value

    def pattern_stats(self):
        """-> (bytes_moved, pattern_columns, family_columns) of the last likelihood call (diagnostic)."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.blg_pattern_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def graph_stats(self):
        """-> (captures, replays): likelihood calls captured as CUDA graphs / replayed from one (diagnostic)."""
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self.L.blg_graph_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_timing(self, on=True):
        self._ck(self.L.blg_set_timing(self.h, 1 if on else 0))

    def last_timing(self):
        a, b = C.c_float(), C.c_float()
        self._ck(self.L.blg_last_timing(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fp64_peak(self):
        """Measured fp64 FMA throughput of this handle's GPU in TFLOP/s (diagnostic, not in the reference)."""
        out = C.c_double()
        self._ck(self.L.blg_fp64_peak(self.h, C.byref(out)))
        return out.value

    def synchronize(self):
        if self.h is not None:
            self._ck(self.L.blg_sync(self.h))


# ---- profile construction, src/model.jl:40-49 ---------------------------------------------------------
def _as_table(df):
    """Accept a pandas DataFrame, a (names, counts) pair or a dict name -> column."""
    def ints(a):                       # integer tables keep their width (10^6 x 100 counts: no 800 MB int64 copy)
        a = np.asarray(a)
        return a if a.dtype.kind in "iu" else a.astype(np.int64)
    if hasattr(df, "columns") and hasattr(df, "to_numpy"):
        return [str(c) for c in df.columns], ints(df.to_numpy())
    if isinstance(df, dict):
        names = list(df.keys())
        return names, np.stack([ints(df[k]) for k in names], axis=1)
    names, counts = df
    return list(names), ints(counts)


def site_patterns(counts, device=None):
    """Unique rows of the count table and their multiplicities, in order of first appearance (what the reference's
    `by(df, names(df), x->size(x)[1])` returns, src/model.jl:41).  On a CUDA device the grouping runs there
    (blg_site_patterns: content-keyed hash table); `device=None` uses it when a device and the library are present
    and the table is large, otherwise numpy."""
    counts = np.asarray(counts)
    N = counts.shape[0]
    use = device is not None
    if device is None and N >= 50_000 and os.path.exists(_lib.SO):
        try:
            import torch
            use = torch.cuda.is_available()
            device = torch.cuda.current_device() if use else None
        except ImportError:
            use = False
    if use and (counts.dtype.itemsize < 4 or counts.dtype == np.int32 or int(counts.max(initial=0)) < 2 ** 31):
        rep = np.zeros(N, dtype=np.int32)
        wgt = np.zeros(N, dtype=np.int64)
        n = C.c_int()
        if counts.dtype == np.uint16:              # (the device sampler's tables: uploaded as they are)
            tab = np.ascontiguousarray(counts)
            rc = _lib.lib().blg_site_patterns_u16(int(device), tab.ctypes.data_as(C.POINTER(C.c_uint16)), N, tab.shape[1],
                                                  rep.ctypes.data_as(C.POINTER(C.c_int32)), wgt.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n))
        else:
            tab = np.ascontiguousarray(counts, dtype=np.int32)
            rc = _lib.lib().blg_site_patterns(int(device), tab.ctypes.data_as(C.POINTER(C.c_int32)), N, tab.shape[1],
                                              rep.ctypes.data_as(C.POINTER(C.c_int32)), wgt.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(n))
        if rc != 0:
            raise _lib.BelugaError(_lib.lib().blg_last_error(None).decode())
        rep = rep[: n.value]
        return tab[rep], wgt[: n.value].copy()
    uniq, first, mult = np.unique(counts, axis=0, return_index=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    return uniq[order], mult[order]


def profile(t, leaves, df, device=None):
    """profile(t, l, df): site-pattern compression (unique rows + multiplicity, first-appearance
    order) and extended counts (leaf = observed count, internal = Σ children).
    -> (X [npatterns × nnodes], weights, m)."""
    names, counts = _as_table(df)
    col = {nm: j for j, nm in enumerate(names)}
    if counts.shape[0] == 0:
        raise ValueError("empty count table")
    uniq, mult = site_patterns(counts, device)
    nn = max(leaves) if leaves else 1
    nodes = []
    stack = [t]
    while stack:
        n = stack.pop()
        nodes.append(n)
        stack.extend(n.c)
    nn = len(nodes)
    # built one node at a time, every node a contiguous run (column-major storage): the matrix handed back is the
    # transposed view, and PArray passes it to the library without another copy
    Mt = np.zeros((nn, uniq.shape[0]), dtype=np.int32)
    ut = np.ascontiguousarray(uniq.T, dtype=np.int32)
    for n in reversed(nodes):          # reversed pre-order visits children before parents
        if not n.c:
            Mt[n.i - 1] = ut[col[leaves[n.i]]]
        else:
            kids = [c.i - 1 for c in n.c]
            np.add(Mt[kids[0]], Mt[kids[1]], out=Mt[n.i - 1]) if len(kids) == 2 else np.sum(Mt[kids], axis=0, out=Mt[n.i - 1])
    return Mt.T, mult.astype(np.int64), int(Mt.max()) + 1


def DLWGD_(nw, df=None, lam=1.0, mu=1.0, eta=0.9, nt=BRANCH, device=0, distributed=False):
    """(model, data) = DLWGD(nw, df, λ, μ, η, nt)   (src/model.jl:22-36).

    With ``distributed=True`` every rank passes the SAME table; patterns are compressed globally and
    each rank keeps a contiguous block of them (families shard across GPUs)."""
    model, t, l = build_model(nw, lam, mu, eta, nt)
    if df is None:
        return model, PArray()
    X, w, _m = profile(t, l, df)
    if distributed:
        rank, world = dist_info()
        lo, hi = shard_range(X.shape[0], rank, world)
        X, w = X[lo:hi], w[lo:hi]
    return model, PArray(X, w, device=device, distributed=distributed)


# ---- free functions with the reference's names ------------------------------------------------------------
def logpdf_(x, p):
    return p.logpdf_(x)


def logpdfroot(n, p):
    return p.logpdfroot(n)


def gradient(d, p):
    """gradient(d::DLWGD, p::PArray)  (src/profile.jl:71-72)"""
    return p.gradient(d)[0]


def gradient_cr(d, p):
    """gradient_cr(d::DLWGD, p::PArray)  (src/profile.jl:74-75)"""
    return p.gradient(d, cr=True)[0]


def set_(x):
    """set!(p::PArray) (profile.jl:79) or set!(d::DLWGD) (model.jl:56-60)."""
    x.set_()


def rev_(p):
    p.rev_()


def extend_(p, i):
    p.extend_(i)


def shrink_(p, i):
    p.shrink_(i)
