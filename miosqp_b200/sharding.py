"""
Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards by independent units:
  * many MIQPs (BASELINE config 2): `shard_range` gives every rank a contiguous block of instances; each rank
    owns those instances' factors and B&B trees -- no data-path collective, results gathered at the end.
  * one big MIQP (config 4): every rank holds the factor and replays the same tree; each B&B step the unsolved
    leaves are dealt round-robin to the ranks (`split_nodes`), solved on the local GPU, and the per-node results
    are exchanged with ONE all-gather; the incumbent is then agreed with ONE all-reduce(MIN)
    (`agree_incumbent`), which also proves that the replicated replays did not diverge.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank` out of `world`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def split_nodes(n_nodes, rank, world):
    """Indices of the frontier batch solved by `rank` (round-robin, so neighbouring siblings land on different GPUs)."""
    return list(range(rank, n_nodes, world))


def exchange_node_results(local, n_nodes, rank, world, group=None):
    """`local`: {node index: (status, iters, seconds, x, y)} solved on this rank -> the full list on every rank."""
    import torch.distributed as dist
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    assert len(merged) == n_nodes
    return [merged[k] for k in range(n_nodes)]


def agree_incumbent(upper_glob, group=None, device=None):
    """all-reduce(MIN) of the incumbent upper bound; returns (global minimum, True if this rank already had it)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([upper_glob if np.isfinite(upper_glob) else 1e300], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    best = float(t.item())
    best = np.inf if best >= 1e300 else best
    return best, (best == upper_glob or (np.isinf(best) and np.isinf(upper_glob)))


def gather_results(local_results, group=None):
    """Concatenate per-rank result lists in rank order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, local_results, group=group)
    return [r for part in out for r in part]
