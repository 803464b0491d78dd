"""Utterance-parallel sharding across the GPUs of one box (one process per GPU, weights replicated).

The reference has no distributed layer at all: its only data parallelism is N independent server workers
(examples/server/server.cpp:225-321).  Utterances are independent, so the path shards with NO data-path collective:
rank 0 deals utterances to ranks (length-balanced), each rank runs its own batched forward, and PCM comes back in the
caller's order.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only to scatter the prompts and gather PCM.
"""
from __future__ import annotations

import numpy as np


def plan_shards(n_tokens: list[int], world: int) -> list[list[int]]:
    """Length-balanced deal: sort by length (desc) and snake across ranks so every rank gets ~equal work and batch size.
    Returns, per rank, the list of original utterance indices."""
    order = sorted(range(len(n_tokens)), key=lambda i: (-n_tokens[i], i))
    shards: list[list[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        rnd, k = divmod(pos, world)
        shards[k if rnd % 2 == 0 else world - 1 - k].append(idx)
    return shards


def scatter_prompts(dist, prompts: list[list[int]] | None, src: int = 0):
    """rank `src` holds all prompts; every rank receives (its utterance indices, its prompts)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if rank == src:
        shards = plan_shards([len(p) for p in prompts], world)
        payload = [(s, [prompts[i] for i in s]) for s in shards]
    else:
        payload = [None] * world
    out = [None]
    dist.scatter_object_list(out, payload, src=src)
    return out[0]


def gather_pcm(dist, idx: list[int], pcms: list[np.ndarray], n_total: int, dst: int = 0, device=None):
    """Gather variable-length PCM to rank `dst`, restoring the caller's utterance order.  Uses tensor collectives
    (NCCL over NVLink on GPUs): first the lengths, then one padded float32 buffer per rank."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device or "cpu"
    lens = torch.tensor([len(p) for p in pcms] + [0] * 0, dtype=torch.int64, device=dev)
    meta = torch.tensor([len(pcms), int(lens.sum()) if len(pcms) else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_n, max_s = int(max(m[0] for m in metas)), int(max(m[1] for m in metas))
    pad_l = torch.zeros(max_n, dtype=torch.int64, device=dev); pad_l[:len(pcms)] = lens
    pad_i = torch.full((max_n,), -1, dtype=torch.int64, device=dev); pad_i[:len(idx)] = torch.tensor(idx, dtype=torch.int64, device=dev)
    flat = torch.zeros(max_s, dtype=torch.float32, device=dev)
    if len(pcms):
        flat[:int(lens.sum())] = torch.from_numpy(np.concatenate(pcms)).to(dev)
    gl = [torch.zeros_like(pad_l) for _ in range(world)] if rank == dst else None
    gi = [torch.zeros_like(pad_i) for _ in range(world)] if rank == dst else None
    gf = [torch.zeros_like(flat) for _ in range(world)] if rank == dst else None
    dist.gather(pad_l, gl, dst=dst)
    dist.gather(pad_i, gi, dst=dst)
    dist.gather(flat, gf, dst=dst)
    if rank != dst:
        return None
    out: list[np.ndarray | None] = [None] * n_total
    for r in range(world):
        off = 0
        buf = gf[r].cpu().numpy()
        for n, i in zip(gl[r].tolist(), gi[r].tolist()):
            if i >= 0:
                out[i] = buf[off:off + n].copy()
            off += n
    return out
