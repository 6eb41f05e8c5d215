"""world_size-2 gloo tests of the multi-GPU decomposition logic on CPU (oracle-backed stand-in projector)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_l2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_projector import OracleProjector
        from whitomo_b200.distributed import AngleSplit, shard_range
        from oracle import white as ow
        rng = np.random.default_rng(0)
        n, na = 24, 21
        P = OracleProjector(n, n, int(np.sqrt(2) * n + 1), np.linspace(0, np.pi, na))
        x = torch.from_numpy(rng.random((n, n)).astype(np.float32))
        y = torch.from_numpy(rng.random(P.sino_shape).astype(np.float32))
        sp = AngleSplit(P)
        a0, a1 = sp.range
        full_fp, full_bp = P.fp(x), P.bp(y)
        loc = sp.fp(x)
        ok = bool(torch.equal(loc[a0:a1], full_fp[a0:a1])) and float(loc[:a0].abs().sum() + loc[a1:].abs().sum()) == 0.0
        ok = ok and rel_l2(sp.gather_sinogram(loc).numpy(), full_fp.numpy()) < 1e-7
        ok = ok and rel_l2(sp.bp(y, scale=-1.0).numpy(), -full_bp.numpy()) < 1e-6      # row blocks + async all-reduces
        one = AngleSplit(P, bp_blocks=1)                                                 # one backprojection, one all-reduce
        ok = ok and rel_l2(one.bp(y, scale=-1.0).numpy(), -full_bp.numpy()) < 1e-6
        # white-beam func/grad split by angle == unsplit
        E = 4
        grid = np.stack([np.linspace(1, 9, E), rng.uniform(0.5, 1.5, E)]).T
        wo = ow.WhiteOracle(rng.uniform(0.2, 2.0, E), 1e-5, rng.uniform(200, 4000, (2, E)), grid, (n, n), geometry=P.geom)
        gt = rng.uniform(0, 1, (2, n, n))
        wo.calc_sinogram_with_gt(gt)
        c = rng.uniform(0.1, 0.3, (2, n, n))
        meas = torch.from_numpy(wo.sinogram.reshape(P.sino_shape).astype(np.float32))
        loss, g = sp.white_func_grad(wo, torch.from_numpy(c.astype(np.float32)), meas)
        ok = ok and abs(loss.item() - wo.func(c)) < 1e-5 * wo.func(c) and rel_l2(g.numpy(), wo.grad(c)) < 1e-5
        # slice sharding: ranges tile the stack exactly, every rank reconstructs only its own slices
        s0, s1 = shard_range(7, rank, world)
        cover = torch.zeros(7)
        cover[s0:s1] = 1
        dist.all_reduce(cover)
        ok = ok and bool((cover == 1).all())
        out[rank] = (ok, (a0, a1))
    finally:
        dist.destroy_process_group()


def test_angle_split_and_slice_shards_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    assert all(v[0] for v in out.values()), dict(out)
    assert out[0][1][1] == out[1][1][0] and out[0][1][0] == 0 and out[1][1][1] == 21


def test_range_helpers():
    from whitomo_b200.distributed import shard_range, angle_ranges
    for n, w in ((2048, 8), (7, 3), (2, 4), (0, 2)):
        r = [shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    r = angle_ranges(4096, 8)
    assert r == [(512 * i, 512 * (i + 1)) for i in range(8)]
    for n, w in ((1800, 8), (50, 4), (5, 8)):
        r = angle_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
    assert all(a % 8 == 0 for a, _ in angle_ranges(1800, 8))


def test_cost_balanced_angle_ranges():
    """cuts fall on tile boundaries, cover the angle list, and give the ranks around 45 degrees (small tiles: the
    forward projector pays per tile) fewer angles"""
    from whitomo_b200.distributed import angle_ranges_by_cost
    first, count = [], []
    a = 0
    for t in range(64):                      # 8-angle tiles, except a band of 4-angle tiles in the middle
        c = 4 if 24 <= t < 40 else 8
        first.append(a)
        count.append(c)
        a += c
    r = angle_ranges_by_cost(first, count, a, 4)
    assert r[0][0] == 0 and r[-1][1] == a and all(r[i][1] == r[i + 1][0] for i in range(3))
    assert all(lo in first or lo == a for lo, _ in r)
    cost = lambda lo, hi: sum(0.55 * (0.12 * 8 + 0.88 * c) + 0.45 * c for f, c in zip(first, count) if lo <= f < hi)
    costs = [cost(*x) for x in r]
    assert max(costs) - min(costs) <= 2 * 8                     # within two tiles of each other
    assert (r[1][1] - r[1][0]) <= (r[0][1] - r[0][0])           # the band of small tiles never gets more angles
    # equal tiles: equal counts
    assert angle_ranges_by_cost(list(range(0, 64, 8)), [8] * 8, 64, 4) == [(0, 16), (16, 32), (32, 48), (48, 64)]
    # more ranks than tiles: trailing ranks get empty ranges, nothing is lost
    r = angle_ranges_by_cost([0, 8], [8, 8], 16, 4)
    assert r[0][0] == 0 and r[-1][1] == 16 and sum(hi - lo for lo, hi in r) == 16
