"""World-size-2 test of the family-sharding path on CPU (gloo): every rank evaluates its contiguous
block of site patterns (here with the oracle standing in for the device kernels), the per-rank
scalars are summed with the same all-reduce helper the product uses, and the total must equal the
single-process value (src/profile.jl:56-57 mapreduce semantics)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import beluga_b200 as B
from beluga_b200.model import build_model
from conftest import read_csv, read_nw
from oracle.oracle import NODE as ONODE, OracleDLWGD


def _worker(rank, world, port, nw, names, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert B.dist_info() == (rank, world)
        # same global pattern compression on every rank, then a contiguous shard (DLWGD_(distributed=True))
        _, t, l = build_model(nw, 0.3, 0.4, 0.8, B.NODE)
        X, w, _ = B.profile(t, l, (names, counts))
        lo, hi = B.shard_range(X.shape[0], rank, world)
        # expand the shard back to rows so the oracle sees exactly these patterns with these weights
        leaf_ids = sorted(l)
        rows = np.repeat(X[lo:hi][:, [i - 1 for i in leaf_ids]], w[lo:hi], axis=0)
        orc = OracleDLWGD(nw, [l[i] for i in leaf_ids], rows, 0.3, 0.4, 0.8, ONODE)
        local = torch.tensor([orc.logpdf()], dtype=torch.float64)
        B.allreduce_sum_(local)
        q.put((rank, lo, hi, float(local[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_logpdf_matches_single_process():
    nw = read_nw("plants1.nw")
    names, counts = read_csv("plants1-100.csv")
    want = OracleDLWGD(nw, names, counts, 0.3, 0.4, 0.8, ONODE).logpdf()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nw, names, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == 0 and res[0][2] == res[1][1]
    assert res[0][3] == res[1][3]
    assert res[0][3] == pytest.approx(want, rel=1e-13)
