"""Host-side model graph: the Python mirror of the part of Beluga.jl that STAYS on the host.

Mirrors ``DLWGD`` / ``TreeNode{Node|Branch}`` (src/model.jl:8-36, src/node.jl:12-134): the graph, the θ
dictionaries and the topology edits (addwgd!/addwgt!/removewgd!/reindex!, src/model.jl:106-183).
What does NOT live here any more: ε, W and every per-branch number (src/node.jl:136-218,
src/utils.jl:8-19) — those are built on the device by ``blg_rebuild``.  ``update!``/``set!`` therefore
only *record* which nodes the reference would have refreshed; the PArray replays that list on the
device before the next likelihood call.

Julia names keep their spelling with ``!`` written as a trailing underscore
(``update!`` → ``update_``, ``addwgd!`` → ``addwgd_`` …).
"""
import copy

import numpy as np

from .newick import readnw

NODE, BRANCH = "Node", "Branch"
KIND_CODE = {"root": 0, "sp": 1, "wgd": 2, "wgt": 3, "wgdafter": 4}
ABSENT = 255


_FLAT_ROW = {"λ": 0, "μ": 1, "q": 2, "t": 3}   # rows of DLWGD._flat


class ModelNode:
    """TreeNode{Node{T}} / TreeNode{Branch{T}} (src/node.jl:12-28): id ``i``, payload θ + kind,
    parent ``p``, ordered children ``c``."""

    __slots__ = ("i", "kind", "theta", "p", "c", "model")

    def __init__(self, i, kind, theta, p=None):
        self.i = i
        self.kind = kind
        self.theta = dict(theta)
        self.p = p
        self.c = []
        self.model = None

    # n[:λ], n[:λ, :μ], n[:t] = v          (src/node.jl:32-34)
    def __getitem__(self, key):
        if isinstance(key, tuple):
            return [self.theta[k] for k in key]
        return self.theta[key]

    def __setitem__(self, key, v):
        v = float(v)
        self.theta[key] = v
        m = self.model
        if m is not None:
            m._param_version += 1
            f = m._flat
            if f is not None and m._flat_topo == m._topo_version:
                r = _FLAT_ROW.get(key)
                if r is not None:
                    f[r, self.i - 1] = v

    def __repr__(self):
        return "ModelNode(%d, %s)" % (self.i, self.kind)


def isleaf(n):
    return len(n.c) == 0


def isroot(n):
    return n.p is None


# kind predicates, src/node.jl:68-77
def iswgd(n):
    return n.kind == "wgd"


def iswgt(n):
    return n.kind == "wgt"


def iswgm(n):
    return iswgd(n) or iswgt(n)


def iswgdafter(n):
    return n.kind == "wgdafter" and n.p.kind == "wgd"


def iswgtafter(n):
    return n.kind == "wgdafter" and n.p.kind == "wgt"


def iswgmafter(n):
    return iswgdafter(n) or iswgtafter(n)


def isawgd(n):
    return iswgd(n) or iswgdafter(n) or iswgt(n) or iswgtafter(n)


def issp(n):
    return n.kind == "sp"


def nonwgdparent(n):  # src/node.jl:79-82
    while isawgd(n):
        n = n.p
    return n


def nonwgdchild(n):  # src/node.jl:84-87
    while isawgd(n):
        n = n.c[0]
    return n


def closestnode(n, t):  # src/node.jl:89-97
    t_ = t - n["t"]
    while t_ > 0:
        t = t_
        n = n.p
        t_ -= n["t"]
    return n, t


def postwalk(n):
    """NewickTree.postwalk: children (in order) before the node."""
    out, stack = [], [(n, 0)]
    while stack:
        node, k = stack.pop()
        if k < len(node.c):
            stack.append((node, k + 1))
            stack.append((node.c[k], 0))
        else:
            out.append(node)
    return out


def prewalk(n):
    out, stack = [], [n]
    while stack:
        node = stack.pop()
        out.append(node)
        stack.extend(reversed(node.c))
    return out


class DLWGD:
    """struct DLWGD (src/model.jl:8-11): ``nodes`` Dict id => node, ``leaves`` Dict id => taxon."""

    def __init__(self, nodes, leaves, nt=BRANCH):
        self.nodes = nodes
        self.leaves = leaves
        self.nt = nt
        for n in nodes.values():
            n.model = self
        # device-mirroring bookkeeping (no counterpart in the reference): version counters plus a
        # bounded log of "the reference would have run set!(n) for these ids, in this order"
        # (None = all nodes in post-order).  Each PArray remembers how far it has replayed the log.
        self._topo_version = 0
        self._param_version = 0
        self._flat = None        # [λ; μ; q; t] × node id, kept current by ModelNode.__setitem__
        self._flat_topo = -1
        self._seq = 0
        self._log = []
        self._log_start = 0      # requests with seq <= _log_start have been trimmed away

    def __len__(self):
        return len(self.nodes)

    def __getitem__(self, key):  # d[i], d[i, :λ]   (src/model.jl:15-16)
        if isinstance(key, tuple):
            return self.nodes[key[0]][key[1]]
        return self.nodes[key]

    def __repr__(self):
        return "DLWGD{Float64,%s}(%d)" % (self.nt, len(self))

    @property
    def root(self):
        return self.nodes[1]

    # --- bookkeeping hooks --------------------------------------------------------------------
    def _touch_topology(self):
        self._topo_version += 1
        self._param_version += 1

    def _request(self, ids):
        self._seq += 1
        self._log.append((self._seq, None if ids is None else list(ids)))
        if len(self._log) > 512:
            drop = self._log[:256]
            self._log = self._log[256:]
            self._log_start = drop[-1][0]

    # --- src/model.jl:20 ------------------------------------------------------------------------
    def ne(self):
        return len(self) - 2 * self.nwgd() - 1

    # --- set!(d), src/model.jl:56-60 ------------------------------------------------------------
    def set_(self):
        self._request(None)

    # --- rates, src/model.jl:69-99 --------------------------------------------------------------
    def setrates_(self, X):
        X = np.asarray(X, dtype=np.float64)
        lam, mu = X[0].tolist(), X[1].tolist()
        nodes = self.nodes
        nn = cnt = 0
        for i in sorted(nodes):
            n = nodes[i]
            if isawgd(n):
                break
            th = n.theta                     # (same effect as n["λ"] = …; n["μ"] = … for every node, in bulk)
            th["λ"] = lam[i - 1]
            th["μ"] = mu[i - 1]
            nn = i
            cnt += 1
        self._param_version += 1
        f = self._flat
        if f is not None and self._flat_topo == self._topo_version:
            if cnt == nn:                    # ids 1..nn are the non-WGD nodes (the reference's invariant, model.jl:69-76)
                f[0, :nn] = X[0, :nn]
                f[1, :nn] = X[1, :nn]
            else:
                self._flat = None
        self.set_()

    def getrates(self):
        X = np.zeros((2, self.ne() + 1))
        for i in sorted(self.nodes):
            n = self.nodes[i]
            if isawgd(n):
                break
            X[0, i - 1] = n["λ"]
            X[1, i - 1] = n["μ"]
        return X

    def getwgds(self):  # src/model.jl:211-225 (descending ids)
        wgds = []
        i = max(self.nodes)
        cont = True
        while cont and i >= 1:
            if i not in self.nodes or iswgmafter(self.nodes[i]):
                pass
            elif iswgm(self.nodes[i]):
                wgds.append(i)
            else:
                cont = False
            i -= 1
        return wgds

    def nwgd(self):
        return len(self.getwgds())

    def getqs(self):  # src/model.jl:93-99
        return np.array([self.nodes[i]["q"] for i in self.getwgds()], dtype=np.float64)

    def asvector(self):  # src/model.jl:541-546
        r = self.getrates()
        return np.concatenate([r[0], r[1], self.getqs(), [self.nodes[1]["η"]]])

    def treelength(self):  # src/utils.jl:38
        return sum(n["t"] for n in self.nodes.values())

    # --- WGD insertion / removal, src/model.jl:106-183 ------------------------------------------
    def addwgd_(self, n, t, q):
        return self._addwgm(n, t, q, "wgd")

    def addwgt_(self, n, t, q):
        return self._addwgm(n, t, q, "wgt")

    def _addwgm(self, n, t, q, kind):
        assert not isroot(n), "Cannot add WGD above root node"
        assert n["t"] - t > 0.0, "Invalid WGD time %g - %g < 0." % (n["t"], t)
        i = max(self.nodes) + 1
        w = ModelNode(i, kind, {"t": n["t"] - t, "q": float(q)}, n.p)
        a = ModelNode(i + 1, "wgdafter", {"t": 0.0}, w)
        return self.insertwgd_(n, w, a)

    def insertwgd_(self, n, w, a):
        """addwgd!(d, n, w, a) (src/model.jl:126-138): also used to put a removed pair back."""
        par = n.p
        par.c = [x for x in par.c if x is not n]
        par.c.append(w)
        w.p = par
        w.c = [a]
        a.p = w
        a.c = [n]
        n.p = a
        n.theta["t"] = n["t"] - w["t"]
        self.nodes[w.i] = w
        self.nodes[a.i] = a
        w.model = a.model = self
        self._touch_topology()          # setabove!(nonwgdchild(n)) is subsumed by the full rebuild
        return w

    def removewgd_(self, n, reindex=True, set=True):
        assert iswgm(n), "Not a WGD/T node %d" % n.i
        assert n.i in self.nodes, "Node not in model!"
        parent = n.p
        after = n.c[0]
        child = after.c[0]
        after.c = []
        n.c = []
        parent.c = [x for x in parent.c if x is not n]
        parent.c.append(child)
        child.p = parent
        child.theta["t"] = child["t"] + n["t"]
        del self.nodes[n.i]
        del self.nodes[n.i + 1]
        if reindex:
            self.reindex_(n.i + 2)
        self._touch_topology()
        return child

    def reindex_(self, i):  # src/model.jl:177-183
        for j in range(i, max(self.nodes) + 1):
            nd = self.nodes[j]
            self.nodes[j - 2] = nd
            nd.i = j - 2
            del self.nodes[j]
        self._touch_topology()

    def removewgds_(self):  # src/model.jl:190-195
        for i in self.getwgds():
            self.removewgd_(self.nodes[i], True, False)
        self.set_()

    # --- (model::DLWGD)(θ), src/model.jl:305-339 ------------------------------------------------
    def __call__(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        n = self.ne() + 1
        lam, mu, q, eta = theta[:n], theta[n : 2 * n], theta[2 * n : -1], theta[-1]
        d = {}

        def rec(x, y):
            if isroot(x):
                x_ = ModelNode(1, "root", {"t": 0.0, "λ": lam[0], "μ": mu[0], "η": eta})
            elif issp(x):
                x_ = ModelNode(x.i, "sp", {"t": x["t"], "λ": lam[x.i - 1], "μ": mu[x.i - 1]}, y)
                y.c.append(x_)
            elif iswgm(x):
                j = x.i % n
                j -= (j - 1) // 2
                x_ = ModelNode(x.i, x.kind, {"t": x["t"], "q": q[j - 1]}, y)
                y.c.append(x_)
            else:
                x_ = ModelNode(x.i, "wgdafter", {"t": 0.0}, y)
                y.c.append(x_)
            d[x_.i] = x_
            for c in x.c:
                rec(c, x_)

        rec(self.nodes[1], None)
        return DLWGD(d, dict(self.leaves), self.nt)

    def copy(self):
        return copy.deepcopy(self)

    # --- flat arrays for the C ABI ----------------------------------------------------------------
    def flat_topology(self):
        n = max(self.nodes)
        parent = np.zeros(n, dtype=np.int32)
        kind = np.full(n, ABSENT, dtype=np.uint8)
        cpos = np.zeros(n, dtype=np.int32)
        for i, nd in self.nodes.items():
            kind[i - 1] = KIND_CODE[nd.kind]
            if nd.p is not None:
                parent[i - 1] = nd.p.i
                cpos[i - 1] = next(k for k, x in enumerate(nd.p.c) if x is nd)
        return n, parent, kind, cpos

    def flat_params(self):
        if self._flat is None or self._flat_topo != self._topo_version:
            f = np.zeros((4, max(self.nodes)))
            for i, nd in self.nodes.items():
                th = nd.theta
                for k, r in _FLAT_ROW.items():
                    f[r, i - 1] = th.get(k, 0.0)
            self._flat, self._flat_topo = f, self._topo_version
        f = self._flat
        return f[0], f[1], f[2], f[3], float(self.nodes[1].theta["η"])


# --- update!, src/node.jl:99-116 ---------------------------------------------------------------------
def update_(n, *args, **kw):
    """update!(n) | update!(n, :θ, v) | update!(n, (λ=…, μ=…))  →  update_(n) | update_(n, "λ", v) |
    update_(n, λ=…, μ=…).  Records set!(c) for the children (Node models, setbelow!) and set!(·) along
    the path n → root (setabove!)."""
    if len(args) == 2:
        n[args[0]] = args[1]
    elif len(args) == 1 and isinstance(args[0], dict):
        for k, v in args[0].items():
            n[k] = v
    elif args:
        raise TypeError("update_: bad arguments")
    for k, v in kw.items():
        n[k] = v
    m = n.model
    ids = []
    if m.nt == NODE:
        ids.extend(c.i for c in n.c)
    x = n
    while x is not None:
        ids.append(x.i)
        x = x.p
    m._request(ids)


def initmodel(t, leaves, eta, lam, mu, nt):
    """src/model.jl:227-244: copy the reference tree into model nodes."""
    d = {1: ModelNode(1, "root", {"t": 0.0, "λ": float(lam), "μ": float(mu), "η": float(eta)})}
    stack = [(t, None)]
    while stack:
        x, y = stack.pop()
        if x.p is None:
            x_ = d[1]
        else:
            x_ = ModelNode(x.i, "sp", {"t": float(x.x), "λ": float(lam), "μ": float(mu)}, y)
            d[x.i] = x_
            y.c.append(x_)
        # push children reversed so they are appended to x_.c in file order
        for c in reversed(x.c):
            stack.append((c, x_))
    return d, dict(leaves)


def build_model(nw, lam=1.0, mu=1.0, eta=0.9, nt=BRANCH):
    t, l = readnw(nw)
    nodes, leaves = initmodel(t, l, eta, lam, mu, nt)
    return DLWGD(nodes, leaves, nt), t, l
