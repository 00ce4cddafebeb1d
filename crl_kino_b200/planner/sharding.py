"""Multi-GPU plumbing: planning queries are independent (each owns its tree; the
obstacle set is read-only), so they are split into contiguous blocks, one per
rank, and nothing is exchanged while planning.  The only collective is the
final gather of per-query results on rank 0 (torch.distributed: NCCL on GPUs,
gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_queries, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_queries, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_sizes(n_queries, world):
    return [shard_range(n_queries, r, world)[1] - shard_range(n_queries, r, world)[0] for r in range(world)]


def gather_results(local, n_queries, dst=0, group=None):
    """Gather per-query result rows from every rank onto `dst`.

    local: dict name -> tensor [n_local, ...] (same trailing shape and dtype on
    all ranks; on the NCCL backend the tensors live on the rank's GPU).
    Returns dict name -> tensor [n_queries, ...] on dst (query order restored),
    None elsewhere.  Ragged shards are padded to the largest shard for the
    collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {k: v for k, v in local.items()}
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_queries, world)
    pad = max(sizes)
    out = {} if rank == dst else None
    for name in sorted(local):
        t = local[name]
        buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
        if rank == dst:
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.gather(buf, parts, dst=dst, group=group)
            out[name] = torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
        else:
            dist.gather(buf, None, dst=dst, group=group)
    return out


def reduce_counters(values, op="sum", group=None):
    """All-reduce a small list of python numbers (float64)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(values)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.cpu()]
