"""Sequence sharding for multi-GPU runs: frames of one sequence depend on each other (sdvl.cc:278-281), so the unit of
distribution is the sequence.  No data-path collective: ranks only agree on the elapsed time (max) and the amount of
work done (sum)."""


def shard_seeds(rank, world, seqs_per_rank):
    """Trajectory seeds owned by `rank`: a contiguous, disjoint block per rank (weak scaling)."""
    return [rank * seqs_per_rank + s for s in range(seqs_per_rank)]


def _reduce(value, op, device):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(value, device="cuda"):
    import torch.distributed as dist
    return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value, device="cuda"):
    import torch.distributed as dist
    return _reduce(value, dist.ReduceOp.SUM, device)
