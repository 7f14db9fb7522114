"""World-size-2 gloo test of the data-parallel sharding + final latent gather (host logic of SURVEY.md 8(e))."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edgestyle_b200.dist import gather_latents, guidance_sweep_units, shard_units


def _fake_denoise(unit_ids, scales):
    # stand-in for the per-unit device work: a deterministic function of (unit id, guidance scale)
    return torch.stack([torch.full((4, 8, 8), float(u)) * s + torch.arange(8.0) for u, s in zip(unit_ids, scales)])


def _worker(rank, world, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = guidance_sweep_units(n_units // 4 + 1, [3.0, 4.5, 6.0, 7.5])[:n_units]
    b, e = shard_units(n_units, rank, world)
    local = _fake_denoise([u for u, _ in units[b:e]], [s for _, s in units[b:e]]) if e > b else torch.zeros(0, 4, 8, 8)
    full = gather_latents(local, n_units, rank, world)
    if rank == 0:
        q.put(full)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [4, 5, 1])
def test_two_rank_gather_matches_single_process(n_units):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    units = guidance_sweep_units(n_units // 4 + 1, [3.0, 4.5, 6.0, 7.5])[:n_units]
    want = _fake_denoise([u for u, _ in units], [s for _, s in units])
    assert torch.equal(full, want)


def test_shard_units_partition():
    for n in range(0, 20):
        for world in (1, 2, 3, 8):
            spans = [shard_units(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)
