"""Multi-GPU plumbing (replicas only): independent sequences are assigned to ranks,
no collective sits on the data path; torch.distributed is used for the barrier and for the
max-over-ranks of the timed region."""
from __future__ import annotations


def assign_streams(rank: int, world: int, seqs_per_rank: int, n_streams: int):
    """Stream index used by each of this rank's sequences: round-robin over the distinct synthetic
    streams, offset by rank so that different GPUs start on different streams."""
    return [(b + rank * seqs_per_rank) % n_streams for b in range(seqs_per_rank)]


def global_sequence_ids(rank: int, world: int, seqs_per_rank: int):
    return [rank * seqs_per_rank + b for b in range(seqs_per_rank)]


def max_over_ranks(value_ms: float, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value_ms
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_frames_per_second(frames_per_rank: int, world: int, max_ms: float) -> float:
    return world * frames_per_rank / (max_ms * 1e-3)
