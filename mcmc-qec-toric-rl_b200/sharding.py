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


def sweep_chunks(points, syndromes_per_point, chunk):
    """Cut every point of a threshold sweep (a dict with at least 'd') into work items of at most `chunk` syndromes.
    -> list of (point_index, n_syndromes, cost) with cost ~ n * d^4 (an STDC sample budget of d^4 per chain)."""
    items = []
    for i, pt in enumerate(points):
        left = int(syndromes_per_point)
        while left > 0:
            n = min(int(chunk(pt) if callable(chunk) else chunk), left)
            items.append((i, n, float(n) * float(pt['d']) ** 4))
            left -= n
    return items


def run_sweep_sharded(points, syndromes_per_point, chunk, decode_chunk, rank=None, world=None, group=None):
    """Threshold sweep of generate_data.py's workload (:53-261), sharded: the (point, chunk) work items are balanced by
    cost over the ranks (shard_by_cost), every rank runs `decode_chunk(point, n, item_index) -> (failures, extra)` on
    its items, and the per-point failure counts (generate_data.py:187-189) -- plus whatever `extra` arrays the decoder
    returns, concatenated per point -- are gathered on the host of every rank.  No collective on the data path.
    -> list of dict(point, syndromes, failures, rate, sigma, extra) in point order."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    items = sweep_chunks(points, syndromes_per_point, chunk)
    mine = shard_by_cost([c for (_, _, c) in items], world)[rank]
    # largest first inside the rank as well: the long d = 21 items do not end up at the tail
    local = []
    for j in sorted(mine, key=lambda j: -items[j][2]):
        i, n, _ = items[j]
        failures, extra = decode_chunk(points[i], n, j)
        local.append((j, i, n, int(failures), extra))
    if world == 1:
        parts = [local]
    else:
        parts = [None] * world
        dist.all_gather_object(parts, local, group=group)
    out = [dict(point=pt, syndromes=0, failures=0, extra=[]) for pt in points]
    for (j, i, n, f, extra) in sorted(sum(parts, []), key=lambda t: t[0]):
        out[i]['syndromes'] += n
        out[i]['failures'] += f
        if extra is not None:
            out[i]['extra'].append(extra)
    for o in out:
        n = max(o['syndromes'], 1)
        o['rate'] = o['failures'] / n
        o['sigma'] = float(np.sqrt(max(o['rate'] * (1 - o['rate']), 0.0) / n))
        o['extra'] = np.concatenate(o['extra'], axis=0) if o['extra'] else None
    return out
