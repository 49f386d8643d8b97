"""Prompt-level data parallelism of the generation entrypoint (reference: ``run.sh:3-16`` launches one process per
GPU with disjoint ``--slice start-end`` ranges, ``entrypoints/generate_images.py:39,185-192``).

One process per GPU, prompt ``i`` goes to rank ``i mod world``; nothing is exchanged on the hot path.  The only
communication is the off-path merge of the per-rank statistics (``global_statistics_*.json`` in the reference,
``generate_images.py:297-309``), done here with one ``all_gather_object``.
"""
from __future__ import annotations

from typing import Dict, List, Sequence


def shard_indices(n_prompts: int, rank: int, world: int) -> List[int]:
    """Prompts handled by ``rank``: i = rank, rank + world, ... (the reference's contiguous ``--slice`` is
    ``slice_indices``)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_prompts, world))


def slice_indices(n_prompts: int, rank: int, world: int) -> List[int]:
    """Contiguous ``--slice start-end`` split as in ``run.sh`` (equal ranges, remainder to the last rank)."""
    per = n_prompts // world
    start = rank * per
    end = n_prompts if rank == world - 1 else start + per
    return list(range(start, end))


def merge_statistics(local: Sequence[Dict], group=None) -> List[Dict]:
    """Gather per-prompt records ``{"prompt", "step_compression", "latency", ...}`` from all ranks, ordered by
    prompt index.  Works with any initialised ``torch.distributed`` backend (NCCL on the GPUs, gloo in tests)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return sorted(local, key=lambda r: r.get("index", 0))
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, list(local), group=group)
    merged = [r for part in bucket for r in part]
    return sorted(merged, key=lambda r: r.get("index", 0))


def summarize(records: Sequence[Dict]) -> Dict:
    """Mean accept length (``step_compression``) and images/s over merged records (generate_images.py:297-303)."""
    n = len(records)
    if n == 0:
        return {"n": 0, "mean_accept_length": 0.0, "images_per_s": 0.0}
    acc = sum(r["step_compression"] for r in records) / n
    wall = max(r.get("rank_wall_s", 0.0) for r in records)
    return {"n": n, "mean_accept_length": acc, "images_per_s": (n / wall) if wall > 0 else 0.0}
