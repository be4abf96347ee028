"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on B200s, gloo in CPU tests).

The path shards by variable — the reference's own striping ``variable % threadCount`` (score_main.cpp:136-139) —
and needs exactly two exchange steps: the input (packed codes, or the p*p Gram) is broadcast from
rank 0, and the per-variable caches are gathered to rank 0, which writes the .pss.  Scoring itself uses no
collective.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def stripe(p: int, rank: int, world: int):
    """variables owned by ``rank`` (score_main.cpp:137)."""
    return [v for v in range(p) if v % world == rank]


def broadcast_tensor(t: torch.Tensor | None, shape, dtype, device, src: int = 0) -> torch.Tensor:
    """rank ``src`` passes the tensor, the others pass None; returns the tensor on every rank (on ``device``)."""
    if dist.get_rank() == src:
        buf = t.to(device).contiguous()
    else:
        buf = torch.empty(shape, dtype=dtype, device=device)
    dist.broadcast(buf, src=src)
    return buf


def gather_caches(local: dict, p: int, words: int, device, dst: int = 0):
    """local: {variable: (masks uint64 [n, words], scores float32 [n])} for the variables this rank owns.
    Returns the same dict for ALL p variables on rank ``dst`` (None elsewhere).  Two collectives: an all_gather of
    the per-variable counts, then a padded gather of the packed (mask words, score bits) payload."""
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(p, dtype=torch.int64, device=device)
    for v, (m, s) in local.items():
        counts[v] = len(s)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    owned = sorted(local)
    rows = int(sum(len(local[v][1]) for v in owned))
    payload = np.zeros((rows, words + 1), dtype=np.int64)
    off = 0
    for v in owned:
        m, s = local[v]
        k = len(s)
        payload[off:off + k, :words] = np.ascontiguousarray(m, dtype=np.uint64).view(np.int64).reshape(k, words)
        payload[off:off + k, words] = np.ascontiguousarray(s, dtype=np.float32).view(np.int32).astype(np.int64)
        off += k
    per_rank = [int(c.sum().item()) for c in all_counts]
    width = max(per_rank) if per_rank else 0
    send = torch.zeros((max(width, 1), words + 1), dtype=torch.int64, device=device)
    if rows:
        send[:rows] = torch.from_numpy(payload).to(device)
    recv = [torch.zeros_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    out = {}
    for r in range(world):
        buf = recv[r].cpu().numpy()
        off = 0
        cr = all_counts[r].cpu().numpy()
        for v in range(p):
            k = int(cr[v])
            if v % world != r:
                continue
            masks = buf[off:off + k, :words].astype(np.int64).view(np.uint64).reshape(k, words)
            scores = buf[off:off + k, words].astype(np.int32).view(np.float32)
            out[v] = (masks.copy(), scores.copy())
            off += k
    return out
