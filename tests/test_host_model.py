"""Host-side logic (model graph mirror, profile construction, ABI surface, sharding) against the
oracle.  CPU only — no compute call goes through the CUDA library here."""
import ctypes
import os
import re

import numpy as np
import pytest

import beluga_b200 as B
from beluga_b200 import _lib
from beluga_b200.model import KIND_CODE, build_model
from conftest import DF1, ROOT, S1, read_csv, read_nw
from oracle.oracle import BRANCH as OBRANCH, NODE as ONODE, OracleDLWGD


def assert_same_graph(model, orc):
    n, parent, kind, cpos = model.flat_topology()
    e = orc.export()
    assert n == orc.maxid
    okind = np.where(e["kind"] < 0, 255, e["kind"]).astype(np.uint8)
    np.testing.assert_array_equal(kind, okind)
    present = kind != 255
    np.testing.assert_array_equal(parent[present], e["parent"][present])
    np.testing.assert_array_equal(cpos[present], e["child_pos"][present])
    lam, mu, q, t, eta = model.flat_params()
    np.testing.assert_array_equal(t[present], e["t"][present])
    sp = present & (kind <= 1)
    np.testing.assert_array_equal(lam[sp], e["lam"][sp])
    np.testing.assert_array_equal(mu[sp], e["mu"][sp])
    wg = present & ((kind == 2) | (kind == 3))
    np.testing.assert_array_equal(q[wg], e["q"][wg])
    assert eta == e["eta"]


@pytest.mark.parametrize("tree,csv", [("plants1.nw", "plants1-100.csv"), ("9dicots.nw", "9dicots-f01-100.csv"),
                                      ("12monocots.nw", "12monocots-f01-250.csv")])
def test_newick_ids_and_profile_match_oracle(tree, csv):
    nw = read_nw(tree)
    names, counts = read_csv(csv)
    model, t, l = build_model(nw, 0.3, 0.4, 0.8, B.NODE)
    orc = OracleDLWGD(nw, names, counts, 0.3, 0.4, 0.8, ONODE)
    assert_same_graph(model, orc)
    X, w, m = B.profile(t, l, (names, counts))
    ow, ox = orc.profile()
    np.testing.assert_array_equal(w, ow)
    np.testing.assert_array_equal(X, ox)
    assert max(3, m) == orc.m
    assert w.sum() == counts.shape[0]


def test_wgd_insert_remove_reindex_bookkeeping():      # test/tests.jl:95-110 on the host mirror
    model, _, _ = build_model(S1, 0.2, 0.3, 1 / 1.5, B.NODE)
    orc = OracleDLWGD(S1, *DF1, 0.2, 0.3, 1 / 1.5, ONODE)
    for (i, t, q) in [(3, 2.0, 0.2), (3, 1.0, 0.3), (2, 1.0, 0.4)]:
        w = B.addwgd_(model, model[i], t, q)
        assert w.i == orc.addwgd(i, t, q)
        assert_same_graph(model, orc)
    B.removewgd_(model, model[8])
    orc.removewgd(8)
    assert model[8]["q"] == 0.3 and model[10]["q"] == 0.4
    assert_same_graph(model, orc)
    B.removewgd_(model, model[10])
    orc.removewgd(10)
    assert model[8]["q"] == 0.3
    assert "q" not in model[9].theta
    assert 10 not in model.nodes and 11 not in model.nodes
    assert_same_graph(model, orc)
    assert model.getwgds() == orc.getwgds()
    np.testing.assert_array_equal(model.asvector(), orc.asvector())
    assert model.ne() == orc.ne and model.nwgd() == orc.nwgd


def test_remove_without_reindex_leaves_gap_then_reinsert():   # rjmcmc.jl:591-620 flow
    model, _, _ = build_model(read_nw("plants1.nw"), 0.5, 0.6, 0.7, B.BRANCH)
    orc = OracleDLWGD(read_nw("plants1.nw"), None, None, 0.5, 0.6, 0.7, OBRANCH)
    w1 = B.addwgd_(model, model[6], 0.3, 0.2); orc.addwgd(6, 0.3, 0.2)
    w2 = B.addwgd_(model, model[9], 0.1, 0.6); orc.addwgd(9, 0.1, 0.6)
    after = w1.c[0]
    child = B.removewgd_(model, w1, False, True)
    assert child.i == orc.removewgd(w1.i, False, True)
    assert_same_graph(model, orc)                 # ids 20, 21 absent, 22, 23 present
    n, parent, kind, cpos = model.flat_topology()
    assert kind[w1.i - 1] == 255 and kind[w1.i] == 255
    model.insertwgd_(child, w1, after)            # the rejected-move path puts the pair back
    assert model[w1.i] is w1 and w1.c[0] is after and after.c[0] is child


def test_model_functor_matches_oracle(plants1c):
    nw, (names, counts) = plants1c
    model, _, _ = build_model(nw, 2.0, 1.0, 0.9, B.BRANCH)
    orc = OracleDLWGD(nw, names, counts, 2.0, 1.0, 0.9, OBRANCH)
    B.addwgd_(model, model[6], 0.01, 0.25); orc.addwgd(6, 0.01, 0.25)
    v = model.asvector()
    np.testing.assert_array_equal(v, orc.asvector())
    rng = np.random.default_rng(1)
    th = rng.random(len(v))
    m2 = model(th)
    orc.set_from_vector(th)
    assert_same_graph(m2, orc)
    np.testing.assert_array_equal(m2.asvector(), th)


def test_update_records_reference_refresh_order():
    for nt, want in ((B.BRANCH, [5, 4, 2, 1]), (B.NODE, [6, 7, 5, 4, 2, 1])):
        model, _, _ = build_model(read_nw("plants1.nw"), 0.5, 0.6, 0.7, nt)
        B.update_(model[5], λ=1.5, μ=1.2)          # README.md:50
        assert model._log[-1][1] == want
        assert model[5]["λ"] == 1.5 and model[5]["μ"] == 1.2
        B.update_(model[1], "η", 0.91)
        assert model[1]["η"] == 0.91


def test_flat_parameter_cache_tracks_updates_and_topology_edits():
    # flat_params() serves the per-proposal θ upload from an array that ModelNode.__setitem__ keeps current; it must
    # equal a fresh walk over the graph after any mix of updates, WGD insertions/removals and reindexing
    nw = read_nw("9dicots.nw")
    model, _, _ = build_model(nw, 1.0, 0.92, 0.66, B.BRANCH)

    def fresh():
        n = max(model.nodes)
        out = np.zeros((4, n))
        for i, nd in model.nodes.items():
            for r, k in enumerate(("λ", "μ", "q", "t")):
                out[r, i - 1] = nd.theta.get(k, 0.0)
        return out

    def check():
        lam, mu, q, t, eta = model.flat_params()
        np.testing.assert_array_equal(np.vstack([lam, mu, q, t]), fresh())
        assert eta == model[1]["η"]

    check()
    B.update_(model[5], λ=1.7, μ=0.3); check()
    model[9]["t"] = 0.123; check()
    w = B.addwgd_(model, model[6], 0.01, 0.25); check()
    w["q"] = 0.9; model[6]["λ"] = 2.5; check()
    w2 = B.addwgt_(model, model[12], 0.02, 0.4); check()
    B.removewgd_(model, w); check()                      # reindexes the WGT pair down by two
    B.update_(model[1], "η", 0.5); check()
    X = model.getrates() * 1.3
    model.setrates_(X); check()
    np.testing.assert_array_equal(model.getrates(), X)
    m2 = model(model.asvector() * 1.1)                   # the functor builds an independent model
    assert m2 is not model and m2._flat is not model._flat
    check()


def test_mock_profile_returns_zero():             # profile.jl:28,31,56; test/rjmcmc.jl:3-24 relies on it
    model, data = B.DLWGD(read_nw("plants1.nw"), None, 1.0, 1.0, 0.9)
    assert len(data) == 0
    assert B.logpdf_(model, data) == 0.0
    assert B.logpdf_(model[5], data) == 0.0
    assert B.logpdfroot(model[1], data) == 0.0
    B.extend_(data, 3); B.set_(data); B.rev_(data)


def test_shared_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    header = open(os.path.join(ROOT, "include", "beluga_b200.h")).read()
    declared = set(re.findall(r"\b(blg_[a-z_A-Z0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.SO)
    for s in declared:
        assert hasattr(L, s), s


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.BelugaError):
        B.DLWGD(S1, DF1, 0.2, 0.3, 0.6, B.NODE)


def test_shard_range_partitions():
    for n in (0, 1, 7, 85, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [B.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _split_top(s):
    """Split at top-level commas (brace/paren aware)."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_glue_and_python_mirror_bind_the_header_as_declared():
    """No `julia` in this image: the glue (julia/BelugaB200.jl) cannot run here, so its ccalls are checked statically
    against include/beluga_b200.h — symbol declared, argument count, and the C type class of every argument — and the
    ctypes argtypes of the Python mirror get the same count check."""
    header = open(os.path.join(ROOT, "include", "beluga_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(?:const\s+char\s*\*|int)\s+(blg_[a-z_A-Z0-9]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S):
        args = [a for a in _split_top(" ".join(m.group(2).split())) if a and a != "void"]
        decl[m.group(1)] = args
    assert len(decl) == len(_lib.SYMBOLS), sorted(set(_lib.SYMBOLS) ^ set(decl))

    def c_class(a):
        a = a.replace("const ", "")
        if "**" in a:
            return "pp"
        if "*" in a:
            return "p"
        return "f" if a.split()[0] == "double" else "i"

    def jl_class(t):
        t = t.strip()
        if t.startswith("Ptr{Ptr{"):
            return "pp"
        if t.startswith("Ptr{") or t == "Cstring":
            return "p"
        return "f" if t in ("Float64", "Cdouble") else "i"

    jl = open(os.path.join(ROOT, "julia", "BelugaB200.jl")).read()
    jl = "\n".join(line.split("#")[0] if not line.lstrip().startswith("#") else "" for line in jl.splitlines())
    seen = set()
    for m in re.finditer(r"ccall\(\(:(blg_[a-z_0-9]+),\s*LIBBLG\),\s*(\w+),\s*\(", jl):
        name = m.group(1)
        assert name in decl, "julia glue calls an undeclared symbol: " + name
        # the argument-type tuple: balanced parentheses from the match's end
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(jl[i], 0)
            i += 1
        types = [t for t in _split_top(jl[m.end():i - 1]) if t]
        want = decl[name]
        assert len(types) == len(want), (name, types, want)
        assert [jl_class(t) for t in types] == [c_class(a) for a in want], (name, types, want)
        seen.add(name)
    # every entry the reference's methods map to is bound by the glue
    for name in ("blg_create", "blg_destroy", "blg_set_profiles", "blg_set_topology", "blg_set_params", "blg_rebuild",
                 "blg_logpdf_full", "blg_logpdf_from", "blg_logpdf_root", "blg_gradient", "blg_gradient_cr", "blg_commit",
                 "blg_revert", "blg_extend", "blg_shrink", "blg_comm_unique_id", "blg_comm_init", "blg_last_error"):
        assert name in seen, name
    # Python mirror: argtypes of every symbol it declares have the header's length
    import __graft_entry__
    __graft_entry__.build()
    L = _lib.lib()
    for name, want in decl.items():
        at = getattr(L, name).argtypes
        if at is not None:
            assert len(at) == len(want), (name, at, want)
