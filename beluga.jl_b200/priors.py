"""The reversible-jump priors of src/priors.jl, evaluated on the host (they are O(#nodes) scalar work on the model
graph, which stays on the host like in the reference): ``logpdf(prior, model)`` for ``IRRevJumpPrior``
(priors.jl:131-172), ``BMRevJumpPrior`` (:43-81) and ``CRRevJumpPrior`` (:212-231), and ``gradient(prior, model)``
(:200-205, which the reference gets from ForwardDiff) in closed form for the IR prior.

Distributions are the reference's defaults, spelled out: ``X₀ = MvNormal([1., 1.])`` is — in the Distributions.jl
version the reference pins (0.23) — the ZERO-mean bivariate normal with standard deviations (1, 1); ``πη = Beta(3, 1)``;
``πq = Beta() = Beta(1, 1)``; ``πK = Geometric(0.5)`` on {0, 1, …}.
"""
import math

import numpy as np

from .model import isroot, iswgd, iswgm, iswgmafter, iswgdafter, nonwgdparent


def _beta_logpdf(x, a, b):
    if not (0.0 <= x <= 1.0):
        return -math.inf
    lb = math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
    la = (a - 1.0) * math.log(x) if a != 1.0 else 0.0
    lc = (b - 1.0) * math.log1p(-x) if b != 1.0 else 0.0
    return la + lc - lb


def _mvn_diag_logpdf(x, mean, sd):
    z = (np.asarray(x) - mean) / sd
    return float(-0.5 * (z * z).sum() - np.log(sd).sum() - len(z) * 0.5 * math.log(2.0 * math.pi))


def parentdist(n, m):
    """parentdist(n, m) (src/utils.jl:28-35): branch length from n up to its ancestor m, through any WGD nodes in between."""
    d = n["t"]
    while n.p is not m:
        n = n.p
        d += n["t"]
    return d


def logp_pics(Psi, J, A, q, n):
    """p(Y | Ψ, q) with the covariance integrated out under an inverse-Wishart(q, Ψ) (priors.jl:83-88)."""
    return math.log(J) + (q / 2.0) * math.log(np.linalg.det(Psi)) - ((q + n) / 2.0) * math.log(np.linalg.det(Psi + A))


class _RevJumpPrior:
    def __init__(self, Tl, Psi, x0_mean=(0.0, 0.0), x0_sd=(1.0, 1.0), eta=(3.0, 1.0), q=(1.0, 1.0), pK=0.5):
        self.Tl = float(Tl)
        self.Psi = np.asarray(Psi, dtype=np.float64)
        assert np.all(np.linalg.eigvalsh(self.Psi) > 0)                       # @assert isposdef(Ψ)
        self.x0_mean = np.asarray(x0_mean, dtype=np.float64)
        self.x0_sd = np.asarray(x0_sd, dtype=np.float64)
        self.eta, self.q, self.pK = tuple(eta), tuple(q), float(pK)

    def _k_logpdf(self, k):                                                  # Geometric(p) on {0, 1, …}
        return math.log(self.pK) + k * math.log1p(-self.pK)

    def _wgd_term(self, n):
        return _beta_logpdf(n["q"], *self.q) - math.log(self.Tl)


class IRRevJumpPrior(_RevJumpPrior):
    """Independent log-rates around the root's, inverse-Wishart covariance integrated out (priors.jl:131-172); for the
    Branch model.  ``pE``: optional callable giving the log-density of the expected number of lineages at the end of a
    branch, exp(t(λ-μ)) (πE, priors.jl:164-167)."""

    def __init__(self, Tl, Psi=((1.0, 0.0), (0.0, 1.0)), pE=None, **kw):
        super().__init__(Tl, Psi, **kw)
        self.pE = pE

    def _parts(self, model):
        root = model.nodes[1]
        X0 = np.log([root["λ"], root["μ"]])
        N = model.ne()
        Y = np.zeros((N, 2))
        p, k = 0.0, 0
        for i, n in model.nodes.items():
            if iswgmafter(n):
                continue
            if iswgm(n):
                p += self._wgd_term(n)
                k += 1
            elif isroot(n):
                p += _beta_logpdf(n["η"], *self.eta) + _mvn_diag_logpdf(X0, self.x0_mean, self.x0_sd)
            else:
                Y[i - 2] = np.log([n["λ"], n["μ"]]) - X0
                if self.pE is not None:
                    t = parentdist(n, nonwgdparent(n.p))
                    p += self.pE(math.exp(t * (n["λ"] - n["μ"])))
        return p, k, X0, Y, N

    def logpdf(self, model):
        p, k, X0, Y, N = self._parts(model)
        if not np.all(np.isfinite(Y)) or not np.all(np.isfinite(X0)):
            return -math.inf
        return p + logp_pics(self.Psi, 1.0, Y.T @ Y, 3, N) + self._k_logpdf(k)

    def gradient(self, model):
        """∇ over asvector(model) = [λ_1..λ_n, μ_1..μ_n, q (getwgds order), η] (priors.jl:200-205), in closed form:
        with G = (Ψ + A)⁻¹, ∂ log det(Ψ + A)/∂Y_i = 2·G·Y_i."""
        assert self.pE is None, "gradient with πE: not implemented"
        p, k, X0, Y, N = self._parts(model)
        n = N + 1
        G = np.linalg.inv(self.Psi + Y.T @ Y)
        dY = -(3 + N) * (Y @ G)                                              # ∂/∂Y_i of -((q+n)/2)·log det(Ψ + A)
        g = np.zeros(2 * n + k + 1)
        rates = np.array([[model.nodes[i]["λ"], model.nodes[i]["μ"]] for i in range(1, n + 1)])
        # non-root nodes: Y_i = log θ_i - X0; the root: every Y_i moves by -1, plus the X₀ prior
        g[1:n] = dY[:, 0] / rates[1:, 0]
        g[n + 1:2 * n] = dY[:, 1] / rates[1:, 1]
        dX0 = -dY.sum(axis=0) - (X0 - self.x0_mean) / self.x0_sd ** 2
        g[0] = dX0[0] / rates[0, 0]
        g[n] = dX0[1] / rates[0, 1]
        a, b = self.q
        for j, w in enumerate(model.getwgds()):
            q = model.nodes[w]["q"]
            g[2 * n + j] = (a - 1.0) / q - (b - 1.0) / (1.0 - q)
        a, b = self.eta
        eta = model.nodes[1]["η"]
        g[-1] = (a - 1.0) / eta - ((b - 1.0) / (1.0 - eta) if b != 1.0 else 0.0)
        return g


class BMRevJumpPrior(_RevJumpPrior):
    """Brownian-motion (autocorrelated) log-rates, for the Node model (priors.jl:43-81).  Like the reference, only
    ``iswgd`` nodes count as WGDs here."""

    def __init__(self, Tl, Psi=((500.0, 0.0), (0.0, 500.0)), **kw):
        super().__init__(Tl, Psi, **kw)

    def logpdf(self, model):
        N = model.ne()
        Y = np.zeros((N, 2))
        p, k, logJ = 0.0, 0, 0.0
        for i, n in model.nodes.items():
            if iswgdafter(n):
                continue
            if iswgd(n):
                p += self._wgd_term(n)
                k += 1
            elif isroot(n):
                p += _beta_logpdf(n["η"], *self.eta) + _mvn_diag_logpdf(np.log([n["λ"], n["μ"]]), self.x0_mean, self.x0_sd)
            else:
                pt = nonwgdparent(n.p)
                dt = parentdist(n, pt)
                Y[i - 2] = (np.log([n["λ"], n["μ"]]) - np.log([pt["λ"], pt["μ"]])) / math.sqrt(dt)
                logJ += math.log(dt)
        # J^(-M/2) with M = 2
        return p + (-logJ) + logp_pics(self.Psi, 1.0, Y.T @ Y, 3, N) + self._k_logpdf(k)


class CRRevJumpPrior(_RevJumpPrior):
    """Constant rates across the tree (priors.jl:212-231)."""

    def __init__(self, Tl, **kw):
        super().__init__(Tl, ((1.0, 0.0), (0.0, 1.0)), **kw)

    def logpdf(self, model):
        root = model.nodes[1]
        p = _beta_logpdf(root["η"], *self.eta) + _mvn_diag_logpdf(np.log([root["λ"], root["μ"]]), self.x0_mean, self.x0_sd)
        k = 0
        for w in model.getwgds():
            p += self._wgd_term(model.nodes[w])
            k += 1
        return p + self._k_logpdf(k)
