"""The caller of the hot path: the move schedule of ``mcmc!`` / ``rjmcmc!`` (src/rjmcmc.jl:275-305), replayed
through the same calls the reference makes per proposal — ``update!`` → ``logpdf!(node, data)`` →
``set!(data)`` / ``rev!(data)``; ``logpdfroot`` for η; ``addwgd!`` + ``extend!`` / ``removewgd!`` + ``shrink!`` for
the reversible jumps.

What is mirrored: the Metropolis-within-Gibbs sweep over the tree (rjmcmc.jl:361-379), move_node! (442-461),
move_root! (463-484), move_wgdtime! (487-521), move_wgdrates! (525-549), move_addwgd!/move_rmwgd! with the
simple kernel (559-620), and the accept/reject bookkeeping.  What is NOT: the priors (src/priors.jl) and the
adaptive proposal kernels (AdaptiveMCMC.jl) are out of scope; this driver uses a flat
prior (optionally a Beta(1,·)/bounded-uniform guard so that parameters stay proper) and fixed-width symmetric
proposals.  It exists to exercise and time the likelihood path exactly the way the sampler does.
"""
import math

import numpy as np

from .model import (isawgd, isroot, iswgm, iswgmafter, nonwgdchild, postwalk, update_)


class RevJumpChain:
    """RevJumpChain (src/rjmcmc.jl:127-137): data, model, state (logp, logπ, k)."""

    def __init__(self, data, model, seed=20261019, step=0.05, max_wgds=20, rate_bounds=(1e-4, 50.0), prior=None):
        self.data = data
        self.model = model
        self.prior = prior              # an object with logpdf(model) (priors.py: IRRevJumpPrior, …); None: flat within the bounds
        self.rng = np.random.default_rng(seed)
        self.step = float(step)
        self.max_wgds = int(max_wgds)
        self.lo, self.hi = rate_bounds
        self.state = {"gen": 0, "logp": -math.inf, "logπ": 0.0, "k": model.nwgd()}
        self.accepted = 0
        self.proposed = 0
        self.evals = 0          # likelihood calls issued (each one is a round trip to the device)
        self.trace = []

    # init!(chain) (rjmcmc.jl:148-155)
    def init_(self):
        self.state["logπ"] = self.logprior()
        self.state["logp"] = self.data.logpdf_(self.model)
        self.evals += 1
        self.data.set_()
        return self

    def logprior(self):
        for n in self.model.nodes.values():
            th = n.theta
            if "λ" in th and not (self.lo < th["λ"] < self.hi and self.lo < th["μ"] < self.hi):
                return -math.inf
            if "q" in th and not (0.0 < th["q"] < 1.0):
                return -math.inf
        eta = self.model.nodes[1]["η"]
        if not (0.0 < eta <= 1.0):
            return -math.inf
        return self.prior.logpdf(self.model) if self.prior is not None else 0.0   # logpdf(chain.prior, model), rjmcmc.jl:150

    # acceptreject_da (rjmcmc.jl:311-320): prior first, the likelihood closure only if that stage accepts
    def _acceptreject(self, f, q=0.0):
        self.proposed += 1
        pi = self.logprior()
        a1 = min(pi - self.state["logπ"] + q, 0.0)
        if not (math.log(self.rng.random()) < a1):
            return False, -math.inf, pi
        l = f()
        self.evals += 1
        a2 = min(l - self.state["logp"], 0.0)
        return bool(math.log(self.rng.random()) < a2), l, pi

    def _commit(self, l, pi):
        self.state["logp"] = l
        self.state["logπ"] = pi
        self.data.set_()
        self.accepted += 1

    # ---- fixed-dimension moves --------------------------------------------------------------------------
    def move_node_(self, n):                                   # rjmcmc.jl:442-461
        lam, mu = n["λ"], n["μ"]
        w = np.log([lam, mu]) + self.rng.normal(0.0, self.step, 2)
        update_(n, λ=float(np.exp(w[0])), μ=float(np.exp(w[1])))
        ok, l, pi = self._acceptreject(lambda: self.data.logpdf_(n))
        if ok:
            self._commit(l, pi)
        else:
            update_(n, λ=lam, μ=mu)
            self.data.rev_()

    def move_root_(self, n):                                   # rjmcmc.jl:463-484
        eta = n["η"]
        eta_ = eta + self.rng.normal(0.0, self.step)
        if not (0.0 < eta_ <= 1.0):
            eta_ = eta                                         # reflect-by-reject keeps the call pattern
        update_(n, "η", eta_)
        ok, l, pi = self._acceptreject(lambda: self.data.logpdfroot(n))
        if ok:
            self._commit(l, pi)
        else:
            update_(n, "η", eta)
            self.data.rev_()

    def move_wgdtime_(self, n):                                # rjmcmc.jl:487-521
        child = n.c[0].c[0]
        nextsp = nonwgdchild(n)
        q, t1, t2 = n["q"], n["t"], child["t"]
        u = self.rng.random()
        if u < 0.66:
            q_ = q + self.rng.normal(0.0, self.step)
            n["q"] = q_ if 0.0 < q_ < 1.0 else q
        elif u > 0.33:
            t = self.rng.random() * (t1 + t2)
            n["t"] = t1 + t2 - t
            child["t"] = t
        update_(nextsp)
        ok, l, pi = self._acceptreject(lambda: self.data.logpdf_(nextsp))
        if ok:
            self._commit(l, pi)
        else:
            n["q"] = q
            n["t"] = t1
            child["t"] = t2
            update_(nextsp)
            self.data.rev_()

    def move_wgdrates_(self, n):                               # rjmcmc.jl:525-549
        q = n["q"]
        child = nonwgdchild(n)
        lam, mu = child["λ"], child["μ"]
        q_ = q + self.rng.normal(0.0, self.step)
        n["q"] = q_ if 0.0 < q_ < 1.0 else q
        w = np.log([lam, mu]) + self.rng.normal(0.0, self.step, 2)
        update_(child, λ=float(np.exp(w[0])), μ=float(np.exp(w[1])))
        ok, l, pi = self._acceptreject(lambda: self.data.logpdf_(child))
        if ok:
            self._commit(l, pi)
        else:
            n["q"] = q
            update_(child, λ=lam, μ=mu)
            self.data.rev_()

    # move!(chain, ::IRRevJumpPrior, ::MWGProposals) (rjmcmc.jl:361-379)
    def move_(self):
        for n in postwalk(self.model.nodes[1]):
            if iswgmafter(n):
                continue
            if iswgm(n):
                self.move_wgdtime_(n)
                self.move_wgdrates_(n)
            elif isroot(n):
                self.move_root_(n)
                self.move_node_(n)
            else:
                self.move_node_(n)

    # ---- reversible jumps, simple kernel (rjmcmc.jl:559-620) -----------------------------------------------
    def _randpos(self):                                         # model.jl:197-206
        ids = sorted(self.model.nodes)
        w = np.array([self.model.nodes[i]["t"] for i in ids], dtype=np.float64)
        i = int(self.rng.choice(ids, p=w / w.sum()))
        n = self.model.nodes[i]
        return n, self.rng.random() * n["t"]

    def move_addwgd_(self):
        if self.model.nwgd() + 1 > self.max_wgds:
            return
        n, t = self._randpos()
        if not (n["t"] - t > 0.0):
            return
        child = nonwgdchild(n)
        q = float(self.rng.beta(1.0, 3.0))
        w = self.model.addwgd_(n, t, q)
        self.data.extend_(n.i)
        ok, l, pi = self._acceptreject(lambda: self.data.logpdf_(child))
        if ok:
            self.state["k"] += 1
            self._commit(l, pi)
        else:
            self.model.removewgd_(w)
            self.data.rev_()

    def move_rmwgd_(self):
        wgds = self.model.getwgds()
        if not wgds:
            return
        w = self.model.nodes[int(self.rng.choice(wgds))]
        after = w.c[0]
        n = nonwgdchild(w)
        wi = w.i
        child = self.model.removewgd_(w, False, True)
        ok, l, pi = self._acceptreject(lambda: self.data.logpdf_(n))
        if ok:
            self.data.shrink_(wi)                              # upon acceptance: shrink and reindex
            self.model.reindex_(wi + 2)
            self.state["k"] -= 1
            self._commit(l, pi)
        else:
            self.model.insertwgd_(child, w, after)
            self.data.rev_()

    # ---- drivers -------------------------------------------------------------------------------------------
    def mcmc_(self, n, trace=1):                               # rjmcmc.jl:294-305
        for i in range(n):
            self.state["gen"] += 1
            self.move_()
            if trace and (i + 1) % trace == 0:
                self.trace.append((self.state["gen"], self.state["logp"], self.state["k"]))
        return self

    def rjmcmc_(self, n, trace=1):                             # rjmcmc.jl:275-292
        for i in range(n):
            self.state["gen"] += 1
            if self.rng.random() < 0.5:
                self.move_rmwgd_()
            else:
                self.move_addwgd_()
            self.move_()
            if trace and (i + 1) % trace == 0:
                self.trace.append((self.state["gen"], self.state["logp"], self.state["k"]))
        return self
