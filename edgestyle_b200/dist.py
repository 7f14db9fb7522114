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


# ------------------------------------------------------------------------------------------------------------------
# Running sharded units through the pipeline mirror (BASELINE config 3 / 4)
# ------------------------------------------------------------------------------------------------------------------
def denoise_units(pipe, units: Sequence[Tuple[int, float]], host: dict, num_inference_steps: int, rank: int, world: int,
                  gather: bool = True, **call_kw) -> torch.Tensor:
    """Denoise this rank's share of `units` = [(image index, guidance scale)] with ONE pipeline call (its units form one
    CFG batch: negative rows first, a per-unit guidance vector on the device) and all-gather the final latents.

    host: dict(latents [n_img, 4, h, w], prompt_embeds / negative_prompt_embeds [n_img, 77, ctx], conds: 6 x [n_img, c0,
    h, w] cached conditioning embeddings).  Replaces the sequential sweep over guidance scales of
    /root/reference/test_text2image_pretrained_openpose.py:326-361 (one pipeline call per scale)."""
    b, e = shard_units(len(units), rank, world)
    mine = list(units[b:e])
    h, w = host["latents"].shape[-2:]
    if mine:
        idx = torch.tensor([u for u, _ in mine])
        out = pipe(image=[c[idx] for c in host["conds"]], prompt_embeds=host["prompt_embeds"][idx],
                   negative_prompt_embeds=host["negative_prompt_embeds"][idx], latents=host["latents"][idx],
                   num_inference_steps=num_inference_steps, guidance_scale=torch.tensor([s for _, s in mine]),
                   output_type="latent", **call_kw)
        local = out.images
    else:  # more ranks than units: this rank only takes part in the gather
        local = torch.zeros(0, host["latents"].shape[1], h, w, device=torch.device("cuda", torch.cuda.current_device()))
    if not gather:
        return local
    return gather_latents(local.contiguous(), len(units), rank, world)


_PAIR_GROUPS = {}  # world size -> process groups of the rank pairs (2g, 2g + 1)


def denoise_split_pairs(multi, units: Sequence[Tuple[int, float]], host: dict, num_inference_steps: int, rank: int,
                        world: int, scheduler=None, use_graph: bool = True) -> torch.Tensor:
    """SURVEY.md 8(e) option (ii): every CFG pair is split over TWO GPUs (rank 2u runs the unconditional row of unit u,
    rank 2u + 1 the conditional row) and the two noise predictions are exchanged once per step (one all-gather of
    [1, 4, h, w] fp32 = 64 KB per rank pair over NVLink); both partners then run the same CFG + DDIM update, so their
    latents stay identical without a second exchange.  Halves the per-image latency when there are more GPUs than
    units (config 3: 4 units on 8 GPUs).  world must equal 2 * len(units).  Returns [n_units, 4, h, w] on every rank."""
    import math

    from . import ops
    from .schedulers import DDIMScheduler

    if world != 2 * len(units):
        raise ValueError(f"split pairs need exactly two ranks per unit: {len(units)} units, world {world}")
    u, cond_row = rank // 2, rank % 2
    img, scale = units[u]
    dev = torch.device("cuda", torch.cuda.current_device())
    h, w = host["latents"].shape[-2:]
    eng = multi.engine(1, h, w, use_graph=use_graph)
    pe = host["prompt_embeds"] if cond_row else host["negative_prompt_embeds"]
    eng.set_prompt(pe[img:img + 1].to(dev))
    eng.set_conditioning([c[img:img + 1].to(dev) for c in host["conds"]])
    sch = scheduler or DDIMScheduler()
    ts = sch.set_timesteps(num_inference_steps)
    lat = host["latents"][img:img + 1].to(dev).float().clone()
    pair = None  # world == 2: the default group is the pair
    if world > 2:  # new_group is collective over the default group: every rank creates every pair group -- ONCE per
        # process (a NCCL communicator per call cost 2 s per denoise at 8 ranks: 119 ms/step instead of 9)
        if world not in _PAIR_GROUPS:
            _PAIR_GROUPS[world] = [dist.new_group(ranks=[2 * g, 2 * g + 1]) for g in range(world // 2)]
        pair = _PAIR_GROUPS[world][u]
    eps2 = torch.empty(2, *lat.shape[1:], device=dev, dtype=torch.float32)
    guidance = torch.tensor([float(scale)], device=dev)
    coef = torch.empty(4, device=dev)
    for t in ts:
        eps = eng.step(lat, float(t), (1.0,) * 6)
        dist.all_gather_into_tensor(eps2, eps.contiguous(), group=pair)  # row 0 = unconditional, row 1 = conditional
        a_t, a_p = sch.coefficients(int(t))
        coef.copy_(torch.tensor([math.sqrt(a_t), math.sqrt(1 - a_t), math.sqrt(a_p), math.sqrt(1 - a_p)]))
        ops.cfg_ddim(eps2, lat, guidance, coef)
    out = [torch.empty_like(lat) for _ in range(world)]
    dist.all_gather(out, lat)
    return torch.cat(out[0::2], dim=0)
