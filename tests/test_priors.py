"""The reversible-jump priors (src/priors.jl:43-231) of the host mirror against independent scipy restatements of the
same formulas, and the closed-form prior gradient against central differences."""
import math

import numpy as np
import pytest
from scipy import stats

import beluga_b200 as B
from beluga_b200.model import build_model, nonwgdparent
from beluga_b200.priors import BMRevJumpPrior, CRRevJumpPrior, IRRevJumpPrior, parentdist
from conftest import read_nw


def _model(nt):
    nw = read_nw("9dicots.nw")
    model, t, l = build_model(nw, 1.0, 0.92, 0.66, nt)
    rng = np.random.default_rng(4)
    X = np.exp(rng.normal(0, 0.5, size=(2, model.ne() + 1)))
    model.setrates_(X)
    model.addwgd_(model[6], 0.01, 0.25)
    model.addwgd_(model[14], 0.02, 0.6)
    return model


def test_ir_prior_matches_an_independent_restatement():
    model = _model(B.BRANCH)
    Tl = model.treelength()
    Psi = np.array([[2.0, 0.3], [0.3, 1.5]])
    pr = IRRevJumpPrior(Tl=Tl, Psi=Psi)
    got = pr.logpdf(model)
    # priors.jl:144-172 spelled out with scipy
    n = model.ne() + 1
    rates = np.array([[model[i]["λ"], model[i]["μ"]] for i in range(1, n + 1)])
    X0 = np.log(rates[0])
    Y = np.log(rates[1:]) - X0
    A = Y.T @ Y
    want = stats.beta(3, 1).logpdf(model[1]["η"]) + stats.multivariate_normal([0, 0], np.eye(2)).logpdf(X0)
    for w in model.getwgds():
        want += stats.beta(1, 1).logpdf(model[w]["q"]) - math.log(Tl)
    want += 1.5 * math.log(np.linalg.det(Psi)) - 0.5 * (3 + n - 1) * math.log(np.linalg.det(Psi + A))
    want += stats.geom(0.5, loc=-1).logpmf(2)
    assert got == pytest.approx(want, rel=1e-12)
    # πE: log-density of the expected lineage count at the end of every branch (priors.jl:164-167)
    pe = stats.lognorm(s=1.0)
    pr2 = IRRevJumpPrior(Tl=Tl, Psi=Psi, pE=pe.logpdf)
    extra = sum(pe.logpdf(math.exp(parentdist(model[i], nonwgdparent(model[i].p)) * (model[i]["λ"] - model[i]["μ"]))) for i in range(2, n + 1))
    assert pr2.logpdf(model) == pytest.approx(want + extra, rel=1e-12)


def test_ir_prior_gradient_matches_central_differences():
    model = _model(B.BRANCH)
    pr = IRRevJumpPrior(Tl=model.treelength(), Psi=[[2.0, 0.3], [0.3, 1.5]], q=(1.5, 2.0))
    g = pr.gradient(model)
    n = model.ne() + 1
    assert len(g) == 2 * n + 2 + 1 == len(model.asvector())
    params = [(i, "λ") for i in range(1, n + 1)] + [(i, "μ") for i in range(1, n + 1)] + [(w, "q") for w in model.getwgds()] + [(1, "η")]
    for j, (i, key) in enumerate(params):
        v0 = model[i][key]
        h = 1e-6 * max(abs(v0), 1e-2)
        model[i][key] = v0 + h; up = pr.logpdf(model)
        model[i][key] = v0 - h; dn = pr.logpdf(model)
        model[i][key] = v0
        assert g[j] == pytest.approx((up - dn) / (2 * h), rel=1e-5, abs=1e-6), (j, i, key)


def test_bm_and_cr_priors():
    model = _model(B.NODE)
    Tl = model.treelength()
    pr = BMRevJumpPrior(Tl=Tl)
    n = model.ne() + 1
    Y, logJ = [], 0.0
    for i in range(2, n + 1):
        pt = nonwgdparent(model[i].p)
        dt = parentdist(model[i], pt)
        Y.append((np.log([model[i]["λ"], model[i]["μ"]]) - np.log([pt["λ"], pt["μ"]])) / math.sqrt(dt))
        logJ += math.log(dt)
    Y = np.array(Y)
    Psi = np.diag([500.0, 500.0])
    want = stats.beta(3, 1).logpdf(model[1]["η"]) + stats.multivariate_normal([0, 0], np.eye(2)).logpdf(np.log([model[1]["λ"], model[1]["μ"]]))
    want += 2 * (0.0 - math.log(Tl)) - logJ + 1.5 * math.log(np.linalg.det(Psi)) - 0.5 * (3 + n - 1) * math.log(np.linalg.det(Psi + Y.T @ Y))
    want += stats.geom(0.5, loc=-1).logpmf(2)
    assert pr.logpdf(model) == pytest.approx(want, rel=1e-12)
    cr = CRRevJumpPrior(Tl=Tl)
    wantc = stats.beta(3, 1).logpdf(model[1]["η"]) + stats.multivariate_normal([0, 0], np.eye(2)).logpdf(np.log([model[1]["λ"], model[1]["μ"]]))
    wantc += -2 * math.log(Tl) + stats.geom(0.5, loc=-1).logpmf(2)
    assert cr.logpdf(model) == pytest.approx(wantc, rel=1e-12)
    # out of support
    model[1]["η"] = 1.5
    assert cr.logpdf(model) == -math.inf
