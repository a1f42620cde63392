"""Host-side sharding logic, including a world_size-2 gloo run on CPU (no GPU involved: the decode function is
a stand-in; what is tested is the partition and the gather)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mcmc_qec_toric_rl_b200 import sharding  # noqa: E402


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (10, 4), (10000, 8), (3, 8)])
def test_shard_range_partitions(n, world):
    blocks = [sharding.shard_range(n, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == n
    for (a, b), (c, d) in zip(blocks, blocks[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1


def test_shard_by_cost_balances_threshold_sweep():
    d = np.array([7, 11, 15, 21] * 6)
    costs = d.astype(float) ** 4
    parts = sharding.shard_by_cost(costs, 8)
    assert sorted(sum(parts, [])) == list(range(len(d)))
    loads = np.array([costs[p].sum() for p in parts])
    assert loads.max() <= 1.35 * loads.mean()


def test_count_failures():
    dist = np.array([[90.0, 10, 0, 0], [10, 20, 60, 10], [25, 25, 25, 25]])
    assert sharding.count_failures(dist, [0, 1, 0]) == 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    codes = list(range(11))                                   # stand-ins for code objects
    fake = lambda block: np.array([[c, 2 * c, rank] for c in block], dtype=np.float64)
    out = sharding.decode_sharded(fake, codes)
    dist.destroy_process_group()
    q.put((rank, out))


def test_decode_sharded_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    for r in range(2):
        out = res[r]
        assert out.shape == (11, 3)
        assert np.array_equal(out[:, 0], np.arange(11)) and np.array_equal(out[:, 1], 2 * np.arange(11))
        assert np.array_equal(out[:, 2], np.array([0] * 6 + [1] * 5))   # rank 0 decoded 6 syndromes, rank 1 five
    assert np.array_equal(res[0], res[1])


def test_sweep_chunks_cover_every_point():
    pts = [dict(d=d, p=p) for d in (7, 21) for p in (0.1, 0.2)]
    items = sharding.sweep_chunks(pts, 1000, lambda pt: 300 if pt['d'] == 7 else 128)
    for i, pt in enumerate(pts):
        sizes = [n for (j, n, c) in items if j == i]
        assert sum(sizes) == 1000 and max(sizes) <= (300 if pt['d'] == 7 else 128)
    assert all(abs(c - n * pts[i]['d'] ** 4) < 1e-6 for (i, n, c) in items)


def _sweep_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pts = [dict(d=d, p=p) for d in (7, 11, 15, 21) for p in (0.10, 0.15, 0.20)]
    seen = []

    def fake(pt, n, item):           # stand-in decoder: "fails" on every (d)-th syndrome of the item, remembers who ran it
        seen.append((pt['d'], n))
        return n // pt['d'], np.full((n, 2), rank, dtype=np.int16)
    out = sharding.run_sweep_sharded(pts, 500, 64, fake)
    cost = sum(n * d ** 4 for d, n in seen)
    dist.destroy_process_group()
    q.put((rank, [(o['syndromes'], o['failures'], o['rate'], o['sigma'], o['extra'].shape, int(o['extra'][:, 0].sum())) for o in out], cost))


def test_run_sweep_sharded_gloo_world2():
    """Config 5's control flow on CPU: two ranks, cost-balanced items, host-side gather of failure counts and arrays."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = {}
    for _ in procs:
        r, rows, cost = q.get(timeout=120)
        res[r] = (rows, cost)
    [p.join(timeout=60) for p in procs]
    assert res[0][0] == res[1][0]                                   # every rank holds the whole gathered curve
    ds = [d for d in (7, 11, 15, 21) for _ in range(3)]
    for d, (n, f, rate, sigma, shape, ranksum) in zip(ds, res[0][0]):
        assert n == 500 and shape == (500, 2)
        assert f == 7 * (64 // d) + (500 - 7 * 64) // d             # 7 items of 64 and one of 52
        assert abs(rate - f / 500) < 1e-12 and abs(sigma - np.sqrt(rate * (1 - rate) / 500)) < 1e-12
    c0, c1 = res[0][1], res[1][1]
    assert abs(c0 - c1) <= 0.1 * (c0 + c1)                          # the two ranks carry about the same cost
    assert 0 < sum(r[5] for r in res[0][0]) < 12 * 500              # both ranks contributed rows
