"""
Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards by independent units:
  * many MIQPs (BASELINE config 2): `shard_range` gives every rank a contiguous block of instances; each rank
    owns those instances' factors and B&B trees -- no data-path collective, results gathered at the end.
  * one big MIQP (config 4): every rank holds the factor and replays the same tree; each B&B step the unsolved
    leaves are dealt round-robin to the ranks (`split_nodes`), solved on the local GPU, and the per-node results
    are exchanged with ONE all_gather_into_tensor of a packed FP64 block; the incumbent is then agreed with ONE all-reduce(MIN)
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


class DistCtx(object):
    """Where a rank stands in the frontier split: (rank, world, process group, device of the exchange tensors).
    Built from the 3- or 4-tuples callers pass as `dist_ctx` (device None = host tensors, the gloo CPU tests)."""

    def __init__(self, rank, world, group=None, device=None):
        self.rank, self.world, self.group, self.device = int(rank), int(world), group, device
        self.exchange_s = 0.0       # wall time spent in the collectives (all-gather of node results + all-reduce of the incumbent)
        self.exchanges = 0
        self.exchange_bytes = 0

    @classmethod
    def of(cls, ctx):
        if ctx is None or isinstance(ctx, cls):
            return ctx
        return cls(*ctx)


def exchange_node_results(local, n_nodes, n, m, ctx):
    """`local`: {node index: (status, iters, seconds, x, y)} solved on this rank -> the full list on every rank, with ONE
    pre-sized all_gather_into_tensor: every rank contributes a [ceil(n_nodes / world), 3 + n + m] FP64 block (status, iters,
    seconds, x, y per node; node k lives in row k // world of rank k % world's block, the round-robin deal of split_nodes).
    No pickling, no host round trip beyond the one copy of the block to the exchange device (NCCL: the rank's GPU)."""
    import time
    import torch
    import torch.distributed as dist
    t0 = time.perf_counter()
    per = -(-n_nodes // ctx.world)
    width = 3 + n + m
    block = np.zeros((per, width))
    for k, (st, it, secs, x, y) in local.items():
        row = block[k // ctx.world]
        row[0], row[1], row[2] = st, it, secs
        row[3:3 + n] = x
        row[3 + n:] = y
    send = torch.from_numpy(block)
    if ctx.device is not None:
        send = send.to(ctx.device, non_blocking=False)
    out = torch.empty((ctx.world * per, width), dtype=torch.float64, device=send.device)     # rank r's block = rows r*per ..
    dist.all_gather_into_tensor(out, send, group=ctx.group)
    got = out.cpu().numpy().reshape(ctx.world, per, width)
    merged = []
    for k in range(n_nodes):
        row = got[k % ctx.world, k // ctx.world]
        merged.append((int(row[0]), int(row[1]), float(row[2]), np.array(row[3:3 + n]), np.array(row[3 + n:])))
    ctx.exchange_s += time.perf_counter() - t0
    ctx.exchanges += 1
    ctx.exchange_bytes += int(out.numel() * 8)
    return merged


def agree_incumbent(upper_glob, group=None, device=None, ctx=None):
    """all-reduce(MIN) of the incumbent upper bound (one FP64 value); returns (global minimum, True if this rank already had it)."""
    import time
    import torch
    import torch.distributed as dist
    t0 = time.perf_counter()
    if ctx is not None:
        group, device = ctx.group, ctx.device
    t = torch.tensor([upper_glob if np.isfinite(upper_glob) else 1e300], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    best = float(t.item())
    best = np.inf if best >= 1e300 else best
    if ctx is not None:
        ctx.exchange_s += time.perf_counter() - t0
    return best, (best == upper_glob or (np.isinf(best) and np.isinf(upper_glob)))


def gather_results(local_results, group=None):
    """Concatenate per-rank result lists in rank order on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [None] * world
    dist.all_gather_object(out, local_results, group=group)
    return [r for part in out for r in part]
