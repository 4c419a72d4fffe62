"""The library's own communicator (include/beluga_b200.h, blg_comm_*): families shard across ranks, every likelihood
and gradient entry returns the sum over the ranks (src/profile.jl:27-37,56-75).  One rank on one GPU always runs; the
two-rank test needs two GPUs."""
import ctypes as C
import os

import numpy as np
import pytest

import beluga_b200 as B
from conftest import read_csv, read_nw
from oracle.oracle import BRANCH as OBRANCH, OracleDLWGD

pytestmark = pytest.mark.gpu


def test_one_rank_communicator_goes_through_nccl():
    nw = read_nw("9dicots.nw")
    names, counts = read_csv("9dicots-f01-25.csv")
    model, data = B.DLWGD(nw, (names, counts), 1.0, 0.92, 0.66, B.BRANCH, device=0)
    uid = np.zeros(128, dtype=np.uint8)
    assert data.L.blg_comm_unique_id(uid.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    assert uid.any()
    data._ck(data.L.blg_comm_init(data.h, 1, 0, uid.ctypes.data_as(C.POINTER(C.c_uint8))))
    assert data.L.blg_comm_size(data.h) == 1
    orc = OracleDLWGD(nw, names, counts, 1.0, 0.92, 0.66, OBRANCH)
    assert B.logpdf_(model, data) == pytest.approx(orc.logpdf(), rel=1e-12)
    B.update_(model[5], λ=1.5, μ=1.2); orc.update(5, lam=1.5, mu=1.2)
    assert B.logpdf_(model[5], data) == pytest.approx(orc.logpdf_from(5), rel=1e-12)
    g, ll = data.gradient(model)
    og = np.asarray(orc.gradient())
    assert np.abs(g - og).max() <= 1e-8 * np.abs(og).max()
    assert ll == pytest.approx(orc.logpdf(), rel=1e-12)


def _worker(rank, world, port, nw, names, counts, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {}
        # (a) the product's sharded path: same table on every rank, global compression, contiguous blocks
        model, data = B.DLWGD(nw, (names, counts), 0.3, 0.4, 0.8, B.NODE, device=rank, distributed=True)
        assert data.world == world
        out["full"] = B.logpdf_(model, data)
        B.update_(model[7], λ=0.5, μ=0.2)
        out["from7"] = B.logpdf_(model[7], data)
        B.update_(model[1], "η", 0.6)
        out["root"] = B.logpdfroot(model[1], data)
        g, ll = data.gradient(model)
        out["grad"] = g
        out["gll"] = ll
        out["n"] = len(data)
        # (b) a rank with no patterns at all
        from beluga_b200.model import build_model
        m2, t2, l2 = build_model(nw, 0.3, 0.4, 0.8, B.NODE)
        X, w, _ = B.profile(t2, l2, (names, counts))
        Xs, ws = (X, w) if rank == 0 else (X[:0], w[:0])
        d2 = B.PArray(Xs, ws, device=rank, distributed=True)
        out["lopsided"] = B.logpdf_(m2, d2)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_likelihood_and_gradient():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from oracle.oracle import NODE as ONODE
    nw = read_nw("plants1.nw")
    names, counts = read_csv("plants1-100.csv")
    orc = OracleDLWGD(nw, names, counts, 0.3, 0.4, 0.8, ONODE)
    want_full = orc.logpdf()
    orc.update(7, lam=0.5, mu=0.2)
    want_7 = orc.logpdf_from(7)
    orc.update(1, eta=0.6)
    want_root = orc.logpdfroot()
    want_g = np.asarray(orc.gradient())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nw, names, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=500) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0]["n"] + res[1]["n"] > 0 and abs(res[0]["n"] - res[1]["n"]) <= 1
    for r in (0, 1):
        assert res[r]["full"] == pytest.approx(want_full, rel=1e-12)
        assert res[r]["from7"] == pytest.approx(want_7, rel=1e-12)
        assert res[r]["root"] == pytest.approx(want_root, rel=1e-12)
        assert res[r]["gll"] == pytest.approx(want_root, rel=1e-12)
        assert np.abs(res[r]["grad"] - want_g).max() <= 1e-8 * np.abs(want_g).max()
        assert res[r]["lopsided"] == pytest.approx(want_full, rel=1e-12)
    assert res[0]["full"] == res[1]["full"]                 # the same all-reduced bits on every rank
