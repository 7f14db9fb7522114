"""Data-parallel sharding of the denoise path: one process per GPU, a full weight replica
each, zero per-step communication.  The unit of work is one image-generation row group (an (image,
guidance-scale) pair with BOTH of its CFG rows, so the CFG combine stays local); after the last step the final
latents are all-gathered once (NCCL over NVLink on GPUs, gloo in the CPU tests).

The reference has no inference-time parallelism (single device, sequential Python loops over guidance scales:
/root/reference/test_text2image_pretrained_openpose.py:326-361); this is the extension BASELINE.json asks for.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_units(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `n_units` owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_latents(local: torch.Tensor, n_units: int, rank: int, world: int) -> torch.Tensor:
    """All-gather per-rank latents [n_local, C, h, w] into [n_units, C, h, w] (ragged shards are padded to the
    largest shard for the collective and trimmed afterwards)."""
    if world == 1:
        return local
    sizes = [shard_units(n_units, r, world) for r in range(world)]
    n_max = max(e - b for b, e in sizes)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: e - b] for o, (b, e) in zip(out, sizes)], dim=0)


def guidance_sweep_units(images: int, scales: Sequence[float]) -> List[Tuple[int, float]]:
    """BASELINE config 3: every (image, guidance scale) pair is an independent unit."""
    return [(i, float(s)) for i in range(images) for s in scales]
