"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on B200s, gloo in CPU tests).

The path shards by variable — the reference stripes ``variable % threadCount`` (score_main.cpp:136-139); here the
variables can also be dealt out by predicted cost (:func:`assign_lpt`), because candidate families differ in size by
orders of magnitude — and needs exactly two exchange steps: the input (packed codes, or the p*p
Gram) is broadcast from rank 0 (or all-gathered when every rank produced a block of columns), and the per-variable
caches are gathered to rank 0, which writes the .pss.  Scoring itself uses no collective.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def stripe(p: int, rank: int, world: int):
    """variables owned by ``rank`` (score_main.cpp:137)."""
    return [v for v in range(p) if v % world == rank]


def family_cost(card, variable: int, neighbors: int, max_parents: int) -> float:
    """Predicted cost of one variable's family: the number of contingency-table cells over all candidate sets,
    r_v * sum_{l <= K} e_l(arities of the candidates) (elementary symmetric polynomials) — what the counting /
    marginalising kernels stream."""
    ar = [int(card[i]) for i in range(len(card)) if i != variable and (neighbors >> i) & 1]
    K = min(max_parents, len(ar))
    e = [1.0] + [0.0] * K
    for a in ar:
        for l in range(K, 0, -1):
            e[l] += e[l - 1] * a
    return float(card[variable]) * sum(e)


def assign_lpt(costs, world: int):
    """Longest-processing-time-first: variables in decreasing cost (ties by index) go to the least loaded rank (ties by
    rank).  Deterministic, so every rank computes the same owner list without communication.  -> owner[v]."""
    order = sorted(range(len(costs)), key=lambda v: (-costs[v], v))
    load = [0.0] * world
    owner = [0] * len(costs)
    for v in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[v] = r
        load[r] += costs[v]
    return owner


def broadcast_tensor(t: torch.Tensor | None, shape, dtype, device, src: int = 0) -> torch.Tensor:
    """rank ``src`` passes the tensor, the others pass None; returns the tensor on every rank (on ``device``)."""
    if dist.get_rank() == src:
        buf = t.to(device).contiguous()
    else:
        buf = torch.empty(shape, dtype=dtype, device=device)
    dist.broadcast(buf, src=src)
    return buf


def gather_caches(local: dict, p: int, words: int, device, dst: int = 0, owner=None):
    """local: {variable: (masks uint64 [n, words], scores float32 [n])} for the variables this rank owns
    (``owner[v]`` = its rank; default: the reference's striping v % world).
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
            if (owner[v] if owner is not None else v % world) != r:
                continue
            masks = buf[off:off + k, :words].astype(np.int64).view(np.uint64).reshape(k, words)
            scores = buf[off:off + k, words].astype(np.int32).view(np.float32)
            out[v] = (masks.copy(), scores.copy())
            off += k
    return out
