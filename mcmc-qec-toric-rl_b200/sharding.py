"""Syndrome sharding across the GPUs of one box.

Syndromes are independent units: rank r decodes a contiguous block and the only communication is the final
gather of the per-syndrome class distributions (and failure counts) on the host side.  No collective sits on the
data path, so the same code runs under NCCL (GPU ranks) or gloo (CPU tests)."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n syndromes for `rank`; block sizes differ by at most one."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs, world):
    """Greedy longest-processing-time assignment of work items with unequal cost (e.g. a threshold sweep over
    several code distances, cost ~ d^4..d^5) -> list of index lists, one per rank."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    out = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        out[r].append(int(i))
        load[r] += costs[i]
    return [sorted(x) for x in out]


def decode_sharded(decode_fn, codes, rank=None, world=None, group=None):
    """Run `decode_fn(list_of_codes) -> array [n, n_eq]` on this rank's block and gather all blocks on every rank.

    With torch.distributed initialised (NCCL or gloo) rank/world default to the process group's; the gather moves
    host arrays (all_gather_object), never device memory."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(len(codes), rank, world)
    local = np.asarray(decode_fn(codes[lo:hi])) if hi > lo else None
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, (lo, local), group=group)
    parts = [p for p in parts if p[1] is not None]
    parts.sort(key=lambda t: t[0])
    return np.concatenate([p[1] for p in parts], axis=0)


def count_failures(distributions, true_classes):
    """generate_data.py:187-189: a decode fails when argmax of the class distribution differs from the true class."""
    return int((np.argmax(distributions, axis=1) != np.asarray(true_classes)).sum())
