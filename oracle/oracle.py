"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so`` (the literal CPU restatement of Beluga.jl's
likelihood path, see ``dlwgd_oracle.hpp``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the product
package never does.  Method names follow the reference calls they restate.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NODE, BRANCH = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "dlwgd_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int64)
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, lp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        for f in ("orc_ok", "orc_nnodes", "orc_maxid", "orc_npatterns", "orc_m", "orc_ncols", "orc_nwgd", "orc_ne", "orc_setall"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_export.argtypes = [C.c_void_p, C.c_int, ip, ip, ip, dp, dp, dp, dp, dp, dp, dp]
        L.orc_get_W.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_get_profile.argtypes = [C.c_void_p, lp, lp]
        L.orc_get_column.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, C.c_int]
        L.orc_logpdf_full.restype = C.c_double
        L.orc_logpdf_full.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_logpdf_from.restype = C.c_double
        L.orc_logpdf_from.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.orc_logpdfroot.restype = C.c_double
        L.orc_logpdfroot.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_logpdf_x.restype = C.c_double
        L.orc_logpdf_x.argtypes = [C.c_void_p, lp, C.c_int]
        L.orc_csuros_miklos_x.argtypes = [C.c_void_p, lp, C.c_int, C.c_int, dp, C.c_int]
        for f in ("orc_set", "orc_rev"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = None
        for f in ("orc_extend", "orc_shrink"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            getattr(L, f).restype = None
        L.orc_update.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_setfield.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_update_node.argtypes = [C.c_void_p, C.c_int]
        L.orc_setrates.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_addwgd.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_addwgt.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.orc_removewgd.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_nonwgdchild.argtypes = [C.c_void_p, C.c_int]
        L.orc_getwgds.argtypes = [C.c_void_p, ip, C.c_int]
        L.orc_asvector.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_set_from_vector.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_gradient.argtypes = [C.c_void_p, dp, C.c_int, C.c_int, C.c_int]
        L.orc_gradient_cr.argtypes = [C.c_void_p, dp, C.c_int]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def read_csv(path):
    """CSV.read of a count table: header = taxon names, rows = families."""
    with open(path) as fh:
        names = fh.readline().strip().split(",")
    counts = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
    return names, counts


class OracleDLWGD:
    """(model, data) = DLWGD(nw, df, λ, μ, η, nt)  — model.jl:22-36, both halves in one handle."""

    def __init__(self, nw, names=None, counts=None, lam=1.0, mu=1.0, eta=0.9, nt=BRANCH, threads=1):
        self.L = lib()
        self.threads = threads
        if counts is None:
            names, counts = [], np.zeros((0, 0), dtype=np.int64)
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self.h = C.c_void_p(self.L.orc_new(nw.encode(), arr, len(names), _lp(counts), counts.shape[0], lam, mu, eta, nt))
        if not self.L.orc_ok(self.h):
            raise RuntimeError(self.error())

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    def error(self):
        return self.L.orc_last_error(self.h).decode()

    def _chk(self, r):
        if r < 0:
            raise RuntimeError(self.error())
        return r

    # sizes
    nnodes = property(lambda s: s.L.orc_nnodes(s.h))
    maxid = property(lambda s: s.L.orc_maxid(s.h))
    npatterns = property(lambda s: s.L.orc_npatterns(s.h))
    m = property(lambda s: s.L.orc_m(s.h))
    ncols = property(lambda s: s.L.orc_ncols(s.h))
    nwgd = property(lambda s: s.L.orc_nwgd(s.h))
    ne = property(lambda s: s.L.orc_ne(s.h))

    def export(self):
        n = self.maxid
        parent, kind, cpos = (np.zeros(n, dtype=np.int32) for _ in range(3))
        t, lam, mu, q, e0, e1 = (np.zeros(n) for _ in range(6))
        eta = C.c_double()
        self._chk(self.L.orc_export(self.h, n, _ip(parent), _ip(kind), _ip(cpos), _dp(t), _dp(lam), _dp(mu), _dp(q), C.byref(eta), _dp(e0), _dp(e1)))
        return dict(parent=parent, kind=kind, child_pos=cpos, t=t, lam=lam, mu=mu, q=q, eta=eta.value, eps0=e0, eps1=e1)

    def W(self, i):
        m = self.m
        out = np.zeros(m * m)
        self._chk(self.L.orc_get_W(self.h, i, _dp(out)))
        return out.reshape(m, m).T.copy()  # [n, j]

    def profile(self):
        npat, nc = self.npatterns, self.ncols
        w = np.zeros(npat, dtype=np.int64)
        x = np.zeros((npat, nc), dtype=np.int64)
        self.L.orc_get_profile(self.h, _lp(w), _lp(x))
        return w, x

    def column(self, fam, node, proposal=True, cap=4096):
        out = np.full(cap, np.nan)
        r = self._chk(self.L.orc_get_column(self.h, fam, node, 1 if proposal else 0, _dp(out), cap))
        return out[:r]

    # likelihood
    def logpdf(self, per_family=False):
        pf = np.zeros(max(self.npatterns, 1))
        v = self.L.orc_logpdf_full(self.h, self.threads, _dp(pf))
        return (v, pf[: self.npatterns]) if per_family else v

    def logpdf_from(self, i, per_family=False):
        pf = np.zeros(max(self.npatterns, 1))
        v = self.L.orc_logpdf_from(self.h, i, self.threads, _dp(pf))
        if np.isnan(v) and self.error():
            raise RuntimeError(self.error())
        return (v, pf[: self.npatterns]) if per_family else v

    def logpdfroot(self, per_family=False):
        pf = np.zeros(max(self.npatterns, 1))
        v = self.L.orc_logpdfroot(self.h, self.threads, _dp(pf))
        return (v, pf[: self.npatterns]) if per_family else v

    def logpdf_x(self, x):
        x = np.ascontiguousarray(x, dtype=np.int64)
        return self.L.orc_logpdf_x(self.h, _lp(x), len(x))

    def csuros_miklos_x(self, x, node=1):
        x = np.ascontiguousarray(x, dtype=np.int64)
        out = np.full(int(x.max()) + 1, np.nan)
        self._chk(self.L.orc_csuros_miklos_x(self.h, _lp(x), len(x), node, _dp(out), len(out)))
        return out

    # profile state
    def set(self):
        self.L.orc_set(self.h)

    def rev(self):
        self.L.orc_rev(self.h)

    def extend(self, i):
        self.L.orc_extend(self.h, i)

    def shrink(self, i):
        self.L.orc_shrink(self.h, i)

    # model edits
    def update(self, i, lam=None, mu=None, eta=None):
        nan = float("nan")
        self._chk(self.L.orc_update(self.h, i, nan if lam is None else lam, nan if mu is None else mu, nan if eta is None else eta))

    def setfield(self, i, field, v):
        self._chk(self.L.orc_setfield(self.h, i, {"t": 0, "lam": 1, "mu": 2, "q": 3, "eta": 4}[field], v))

    def update_node(self, i):
        self._chk(self.L.orc_update_node(self.h, i))

    def setall(self):
        self._chk(self.L.orc_setall(self.h))

    def setrates(self, X):
        X = np.asfortranarray(X, dtype=np.float64)
        self._chk(self.L.orc_setrates(self.h, _dp(X), X.shape[1]))

    def addwgd(self, i, t, q):
        return self._chk(self.L.orc_addwgd(self.h, i, t, q))

    def addwgt(self, i, t, q):
        return self._chk(self.L.orc_addwgt(self.h, i, t, q))

    def removewgd(self, i, reindex=True, set=True):
        return self._chk(self.L.orc_removewgd(self.h, i, int(reindex), int(set)))

    def nonwgdchild(self, i):
        return self._chk(self.L.orc_nonwgdchild(self.h, i))

    def getwgds(self):
        out = np.zeros(self.maxid, dtype=np.int32)
        k = self.L.orc_getwgds(self.h, _ip(out), len(out))
        return out[:k].tolist()

    def asvector(self):
        out = np.zeros(2 * self.maxid + 2)
        k = self.L.orc_asvector(self.h, _dp(out), len(out))
        return out[:k].copy()

    def set_from_vector(self, th):
        th = np.ascontiguousarray(th, dtype=np.float64)
        self._chk(self.L.orc_set_from_vector(self.h, _dp(th), len(th)))

    def gradient(self, literal=True):
        out = np.zeros(2 * self.maxid + 2)
        k = self._chk(self.L.orc_gradient(self.h, _dp(out), len(out), int(literal), self.threads))
        return out[:k].copy()

    def gradient_cr(self):
        out = np.zeros(2)
        self._chk(self.L.orc_gradient_cr(self.h, _dp(out), self.threads))
        return out


def max_threads():
    return lib().orc_max_threads()
