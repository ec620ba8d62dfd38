"""Tree-batch data parallelism: contiguous shards of trees per rank and the one collective on the path,
the gather of per-tree result rows.  Plumbing only (torch.distributed; NCCL on GPUs,
gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_trees, rank, world):
    """Contiguous block of ceil(T/G) trees per rank (last ranks may be short or empty)."""
    per = (n_trees + world - 1) // world
    lo = min(n_trees, rank * per)
    hi = min(n_trees, lo + per)
    return lo, hi


def gather_rows(local_rows, n_trees_total=None, group=None):
    """All-gather [T_local, K] result rows into [T_total, K] on every rank.  Equal shards use one
    all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_rows
    t_local = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=local_rows.device)
    counts = [torch.zeros_like(t_local) for _ in range(world)]
    dist.all_gather(counts, t_local, group=group)
    counts = [int(c.item()) for c in counts]
    k = local_rows.shape[1]
    tmax = max(counts)
    if all(c == tmax for c in counts) and dist.get_backend(group) == "nccl":
        out = torch.empty((world * tmax, k), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    pad = torch.zeros((tmax, k), dtype=local_rows.dtype, device=local_rows.device)
    pad[: local_rows.shape[0]] = local_rows
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    if n_trees_total is not None:
        assert out.shape[0] == n_trees_total
    return out
