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


# ---- the device-side forms used inside bench.py's timed strong-scaling step (NCCL; every tensor below lives on the rank's GPU)
class _DevView:
    """a [rows][stride] float32 block of device memory owned by the library, as something torch.as_tensor understands (CUDA array interface)"""

    def __init__(self, ptr: int, rows: int, stride: int):
        self.__cuda_array_interface__ = {"shape": (rows, stride), "typestr": "<f4", "data": (ptr, False), "version": 2}


def scatter_tokens_nccl(dist, torch, device, prompts, counts, src: int = 0):
    """rank `src` holds all prompts (lists of token ids); every rank gets its contiguous share (counts[r] utterances) back as host lists.
    One broadcast of the packed batch (lengths + ids, a few KB) over NCCL + local slicing."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_total = int(sum(counts))
    hdr = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        lens = np.asarray([len(p) for p in prompts], np.int64)
        hdr[0], hdr[1] = n_total, int(lens.sum())
    dist.broadcast(hdr, src=src)
    packed = torch.empty(int(hdr[0]) + int(hdr[1]), dtype=torch.int64, device=device)
    if rank == src:
        flat = np.concatenate([lens] + [np.asarray(p, np.int64) for p in prompts])
        src_t = torch.from_numpy(flat)
        packed.copy_(src_t.pin_memory() if getattr(device, "type", str(device)).startswith("cuda") else src_t, non_blocking=False)
    dist.broadcast(packed, src=src)
    h = packed.cpu().numpy()
    lens, ids = h[:n_total], h[n_total:]
    offs = np.concatenate([[0], np.cumsum(lens)])
    first = int(sum(counts[:rank]))
    return [ids[offs[i]:offs[i + 1]].astype(np.uint32).tolist() for i in range(first, first + counts[rank])]


def device_block(torch, device, ptr: int, rows: int, stride: int):
    """the library's [rows][stride] PCM block in device memory as a torch tensor (no copy)"""
    return torch.as_tensor(_DevView(ptr, rows, stride), device=device)


def gather_pcm_nccl(dist, torch, device, block, n_samples, counts, dst: int = 0, host_out=None):
    """every rank's PCM (rows of `block`, a [utterances][stride] tensor on the rank's device, n_samples[b] valid floats each) to rank `dst`: lengths by all_gather, then
    one exact-length send per rank straight from device memory (grouped irecv on dst), then ONE device -> pinned-host copy on `dst`.
    -> (host float32 tensor, per-utterance lengths in caller order) on dst, else None.  (gloo / CPU tensors work too: the world-size-2 test in tests/test_host_cpu.py)"""
    world, rank = dist.get_world_size(), dist.get_rank()
    mx = max(counts)
    mine = torch.zeros(mx, dtype=torch.int64, device=device)
    if n_samples:
        mine[:len(n_samples)] = torch.tensor(n_samples, dtype=torch.int64, device=device)
    all_l = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(all_l, mine)
    lens = torch.stack(all_l).cpu().numpy()                    # [world][mx]
    totals = lens.sum(axis=1)
    packed = None
    if n_samples:
        packed = torch.cat([block[b, :n] for b, n in enumerate(n_samples)])      # the valid samples, contiguous, still on the device
    if rank != dst:
        if packed is not None and packed.numel():
            dist.send(packed, dst=dst)
        return None
    out = torch.empty(int(totals.sum()), dtype=torch.float32, device=device)
    offs = np.concatenate([[0], np.cumsum(totals)])
    ops = []
    for r in range(world):
        seg = out[int(offs[r]):int(offs[r + 1])]
        if r == dst:
            if packed is not None:
                seg.copy_(packed)
        elif seg.numel():
            ops.append(dist.P2POp(dist.irecv, seg, r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    cuda = getattr(device, "type", str(device)).startswith("cuda")
    if host_out is None or host_out.numel() < out.numel():
        host_out = torch.empty(out.numel(), dtype=torch.float32)
        if cuda:
            host_out = host_out.pin_memory()
    host_out[:out.numel()].copy_(out, non_blocking=cuda)
    if cuda:
        torch.cuda.synchronize()
    per_utt = [int(lens[r, i]) for r in range(world) for i in range(counts[r])]
    return host_out[:out.numel()], per_utt
