"""Pins the CPU oracle to every golden vector / known-answer test the reference holds for the
likelihood path.  CPU only."""
import numpy as np
import pytest

from conftest import (DF1, ETA3, GRAD_GOLDEN, LAM3, MU3, S1, SHOULDBE1, SHOULDBE2, SHOULDBE3)
from oracle.oracle import BRANCH, NODE, OracleDLWGD


def test_dl_model_csuros_miklos():            # test/tests.jl:7-13
    d = OracleDLWGD(S1, *DF1, 0.2, 0.3, 1 / 1.5, NODE)
    _, x = d.profile()
    L = d.csuros_miklos_x(x[0])
    assert L[0] == -np.inf
    np.testing.assert_allclose(L[1:], SHOULDBE1[1:], atol=1e-4, rtol=0)


def test_dl_wgd_model():                      # test/tests.jl:16-26
    d = OracleDLWGD(S1, *DF1, 0.2, 0.3, 1 / 1.5, NODE)
    _, x = d.profile()
    d.addwgd(3, 2.0, 0.5)
    x_ = np.concatenate([x[0], [x[0][2], x[0][2]]])
    L = d.csuros_miklos_x(x_)
    np.testing.assert_allclose(L[1:], SHOULDBE2[1:], atol=1e-4, rtol=0)


def test_extend_remove_partial():             # test/tests.jl:29-52, through the Profile API
    d = OracleDLWGD(S1, *DF1, 0.2, 0.3, 1 / 1.5, NODE)
    d.logpdf()
    np.testing.assert_allclose(d.column(0, 1)[1:], SHOULDBE1[1:], atol=1e-4, rtol=0)
    w = d.addwgd(3, 2.0, 0.5)
    d.extend(3)
    d.logpdf_from(3)
    np.testing.assert_allclose(d.column(0, 1)[1:], SHOULDBE2[1:], atol=1e-4, rtol=0)
    child = d.removewgd(w)
    d.logpdf_from(child)                       # stale columns 8,9 are simply unused (tests.jl:47-48)
    np.testing.assert_allclose(d.column(0, 1)[1:], SHOULDBE1[1:], atol=1e-4, rtol=0)


def test_partial_recomputation_1():           # test/tests.jl:55-72
    d = OracleDLWGD(S1, *DF1, 0.2, 0.3, 1 / 1.5, NODE)
    d.update(3, lam=0.9)
    d.logpdf()
    c = d.column(0, 1)
    assert np.all((c[1:] != np.array(SHOULDBE1[1:])))
    d.update(3, lam=0.2)
    d.logpdf_from(3)
    np.testing.assert_allclose(d.column(0, 1)[1:], SHOULDBE1[1:], atol=1e-4, rtol=0)


def test_partial_recomputation_2(plants1):    # test/tests.jl:75-92 (bit-identical after restore)
    nw, (names, counts) = plants1
    d = OracleDLWGD(nw, names, counts, 0.2, 0.3, 1 / 1.5, NODE)
    l1 = d.logpdf()
    d.update(5, lam=0.67)
    l2 = d.logpdf()
    d.update(3, mu=0.12)
    l3 = d.logpdf()
    d.update(1, eta=0.4, lam=0.1)
    l4 = d.logpdf()
    assert len({l1, l2, l3, l4}) == 4
    d.update(5, lam=0.2)
    d.update(3, mu=0.3)
    d.update(1, eta=1 / 1.5, lam=0.2)
    assert d.logpdf() == l1


def test_insert_remove_wgds():                # test/tests.jl:95-110
    d = OracleDLWGD(S1, *DF1, 0.2, 0.3, 1 / 1.5, NODE)
    _, x = d.profile()
    d.addwgd(3, 2.0, 0.2)
    d.addwgd(3, 1.0, 0.3)
    d.addwgd(2, 1.0, 0.4)
    d.removewgd(8)
    e = d.export()
    assert e["q"][7] == 0.3 and e["q"][9] == 0.4
    d.removewgd(10)
    e = d.export()
    assert e["q"][7] == 0.3
    assert e["kind"][8] == 4                   # node 9 is a wgdafter: no :q
    assert d.maxid == 9
    xx = x[0]
    l = d.logpdf_x(np.concatenate([xx, [xx[2], xx[2]]]))
    assert l == pytest.approx(-9.159735036143305, rel=1e-13)


def test_stress_does_not_throw():             # test/tests.jl:113-128
    rng = np.random.default_rng(333)
    for i in range(200):
        lam = np.exp(rng.normal(0, 5))
        mu = lam if i >= 100 else np.exp(rng.normal(0, 5))
        eta = rng.beta(3, 1 / 3)
        d = OracleDLWGD(S1, *DF1, lam, mu, eta, NODE)
        l = d.logpdf()
        assert not np.isnan(l) and l <= 0.0 or l == -np.inf


def test_profile_full_loglikelihood(plants1):  # test/tests.jl:131-141  (≈ is rtol 1.5e-8; we get 4e-16)
    nw, (names, counts) = plants1
    for lam, mu, eta, want in zip(LAM3, MU3, ETA3, SHOULDBE3):
        d = OracleDLWGD(nw, names, counts, lam, mu, eta, NODE)
        assert d.logpdf() == pytest.approx(want, rel=2e-15)


def test_extending_and_shrinking(plants1):    # test/tests.jl:144-173
    nw, (names, counts) = plants1
    rng = np.random.default_rng(7)
    for lam, mu, eta, want in list(zip(LAM3, MU3, ETA3, SHOULDBE3))[:4]:
        d = OracleDLWGD(nw, names, counts, lam, mu, eta, NODE)
        d.logpdf()
        r = list(rng.integers(2, d.nnodes + 1, size=10))
        wnodes = []
        while len(wnodes) != len(r) or len(r) != 0:
            u = rng.random()
            if len(r) > 0 and u < 0.5:
                j = int(r.pop())
                tj = d.export()["t"][j - 1]
                w = d.addwgd(j, rng.random() * tj, rng.random())
                d.extend(j)
                d.logpdf_from(j)
                wnodes.append(w)
            elif len(wnodes) > 0 and u > 0.5:
                w = wnodes.pop()
                child = d.removewgd(w)
                d.shrink(w)
                d.logpdf_from(child)
        assert d.logpdf() == pytest.approx(want, rel=1e-12)


def test_gradient_golden(plants1c):           # test/gradient.jl:1-13
    nw, (names, counts) = plants1c
    d = OracleDLWGD(nw, names, counts, 2.0, 1.0, 0.9, BRANCH, threads=4)
    g = d.gradient()
    np.testing.assert_allclose(g, GRAD_GOLDEN, atol=1e-4, rtol=0)
    # the dual-number gradient agrees with central differences of the oracle's own logpdf
    v = d.asvector()
    for j in (1, 5, 20, 25, 38):
        hstep = 1e-6 * max(1.0, abs(v[j]))
        vp, vm = v.copy(), v.copy()
        vp[j] += hstep
        vm[j] -= hstep
        dp = OracleDLWGD(nw, names, counts, 2.0, 1.0, 0.9, BRANCH)
        dp.set_from_vector(vp)
        dm = OracleDLWGD(nw, names, counts, 2.0, 1.0, 0.9, BRANCH)
        dm.set_from_vector(vm)
        fd = (dp.logpdf() - dm.logpdf()) / (2 * hstep)
        assert g[j] == pytest.approx(fd, rel=2e-6, abs=1e-6)


def test_readme_check_values(dicots):         # SURVEY.md §8c derived values (restatement, not reference-published)
    nw, (names, counts) = dicots
    d = OracleDLWGD(nw, names, counts, 1.0, 0.92, 0.66, BRANCH)
    assert d.logpdf() == pytest.approx(-254.729304236332, rel=1e-13)
    d.addwgd(6, 0.01, 0.25)
    d.extend(6)
    assert d.logpdf() == pytest.approx(-254.876683590601, rel=1e-13)
