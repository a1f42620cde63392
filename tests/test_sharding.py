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
