"""Multi-GPU plumbing: one process per GPU over torch.distributed (NCCL on B200s, gloo in CPU tests).

The path shards by variable — the reference stripes ``variable % threadCount`` (score_main.cpp:136-139); here the
variables can also be dealt out by predicted cost (:func:`assign_lpt`), because candidate families differ in size by
orders of magnitude — or by (variable, parent-set range) (:func:`plan_ranges`, :func:`exchange_ranges`): a family's
canonical order numbers its sets, any contiguous range of that numbering is scored independently
(``urlgpu_score_range``), and the raw scores travel to the variable's owner, which applies the filters that need the
whole family (``urlgpu_result_from_scores``).  Exchange steps: the input (packed codes, or the p*p
Gram) is broadcast from rank 0 (or all-gathered when every rank produced a block of columns); range shards add one
all-to-all of float32 score arrays; the per-variable caches are gathered to rank 0, which writes the .pss.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


_PINNED = {}


class PeerMemoryUnavailable(RuntimeError):
    """CUDA IPC / peer access could not be set up on some rank.  Raised on EVERY rank together (the ranks agree through a
    collective before anyone proceeds), so callers can fall back to the NCCL forms in step."""


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


def _landing(key: str, numel: int, dtype, pinned: bool) -> torch.Tensor:
    """a host buffer kept between calls (page-locked next to a GPU: pageable device->host copies run at a fraction of PCIe speed)"""
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < numel or buf.dtype != dtype:
        buf = torch.empty(max(numel, 1), dtype=dtype, pin_memory=pinned)
        _PINNED[key] = buf
    return buf[:numel]


def gather_caches(local: dict, p: int, words: int, device, dst: int = 0, owner=None, copy: bool = True):
    """local: {variable: (masks uint64 [n, words], scores float32 [n])} for the variables this rank owns
    (``owner[v]`` = its rank; default: the reference's striping v % world).
    Returns the same dict for ALL p variables on rank ``dst`` (None elsewhere).  Three collectives: an all_gather of
    the per-variable counts, then one padded gather of the mask words and one of the scores (kept apart so that every
    variable's block is a contiguous slice on both sides).  ``copy=False`` returns views into the landing buffers, valid
    until the next call — what a writer that formats the blocks straight away wants."""
    world, rank = dist.get_world_size(), dist.get_rank()
    on_gpu = torch.device(device).type == "cuda"
    counts = torch.zeros(p, dtype=torch.int64)
    owned = sorted(local)
    for v in owned:
        counts[v] = len(local[v][1])
    counts = counts.to(device)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = torch.stack(all_counts).cpu().numpy()          # [world, p]
    per_rank = all_counts.sum(axis=1)
    rows, width = int(per_rank[rank]), max(int(per_rank.max()), 1)
    hm = _landing("send_m", width * words, torch.int64, on_gpu).view(width, words)
    hs = _landing("send_s", width, torch.float32, on_gpu)
    hm_np, hs_np = hm.numpy(), hs.numpy()
    off = 0
    for v in owned:
        m, s = local[v]
        k = len(s)
        hm_np[off:off + k] = np.asarray(m).reshape(k, words).view(np.int64)
        hs_np[off:off + k] = s
        off += k
    send_m = torch.empty((width, words), dtype=torch.int64, device=device)
    send_s = torch.empty(width, dtype=torch.float32, device=device)
    send_m[:rows].copy_(hm[:rows], non_blocking=True)
    send_s[:rows].copy_(hs[:rows], non_blocking=True)
    recv_m = [torch.empty_like(send_m) for _ in range(world)] if rank == dst else None
    recv_s = [torch.empty_like(send_s) for _ in range(world)] if rank == dst else None
    dist.gather(send_m, recv_m, dst=dst)
    dist.gather(send_s, recv_s, dst=dst)
    if rank != dst:
        if on_gpu:
            torch.cuda.current_stream().synchronize()   # the landing buffers are reused by the next call
        return None
    total = int(per_rank.sum())
    land_m = _landing("recv_m", max(total, 1) * words, torch.int64, on_gpu).view(-1, words)
    land_s = _landing("recv_s", max(total, 1), torch.float32, on_gpu)
    base = np.concatenate([[0], np.cumsum(per_rank)]).astype(np.int64)
    for r in range(world):   # only the filled rows of every rank's padded block come to the host
        k = int(per_rank[r])
        if k:
            land_m[int(base[r]):int(base[r]) + k].copy_(recv_m[r][:k], non_blocking=True)
            land_s[int(base[r]):int(base[r]) + k].copy_(recv_s[r][:k], non_blocking=True)
    if on_gpu:
        torch.cuda.current_stream().synchronize()
    lm, ls = land_m.numpy().view(np.uint64), land_s.numpy()
    out = {}
    for r in range(world):
        off = int(base[r])
        for v in range(p):
            if (owner[v] if owner is not None else v % world) != r:
                continue
            k = int(all_counts[r, v])
            masks, scores = lm[off:off + k], ls[off:off + k]
            out[v] = (masks.copy(), scores.copy()) if copy else (masks, scores)
            off += k
    return out


# ---------------------------------------------------------------------------------------------------------------
# (variable, parent-set range) shards

def plan_ranges(sizes, world: int, owner=None, min_chunk: int = 1 << 16):
    """Deal the concatenated index space of all families (sizes[v] = sets of variable v, in variable order) into `world`
    contiguous pieces of (almost) equal size.  -> (pieces, owner): pieces[r] = [(variable, first, count), ...] scored by
    rank r; owner[v] = the rank that assembles variable v's family and filters it (default: the rank scoring the largest
    part of it, ties to the lower rank).  Deterministic: every rank computes the same plan without communication.
    Pieces shorter than `min_chunk` sets are not split off a family's end (a launch per few sets costs more than it saves)."""
    total = int(sum(sizes))
    pieces = [[] for _ in range(world)]
    share = [dict() for _ in range(len(sizes))]
    bounds = [(total * r) // world for r in range(world + 1)]
    # snap the cut points to family ends when the remainder would be tiny
    starts = np.concatenate([[0], np.cumsum(np.asarray(sizes, dtype=np.int64))])
    for r in range(1, world):
        v = int(np.searchsorted(starts, bounds[r], side="right") - 1)
        if v < len(sizes):
            if bounds[r] - starts[v] < min_chunk:
                bounds[r] = int(starts[v])
            elif starts[v + 1] - bounds[r] < min_chunk:
                bounds[r] = int(starts[v + 1])
    for r in range(world):
        lo, hi = max(bounds[r], bounds[r - 1] if r else 0), bounds[r + 1]
        bounds[r] = lo
        v = int(np.searchsorted(starts, lo, side="right") - 1)
        while lo < hi and v < len(sizes):
            end = min(hi, int(starts[v + 1]))
            if end > lo:
                pieces[r].append((v, lo - int(starts[v]), end - lo))
                share[v][r] = share[v].get(r, 0) + end - lo
            lo = end
            v += 1
    if owner is None:
        owner = [min(sh, key=lambda r: (-sh[r], r)) if sh else v % world for v, sh in enumerate(share)]
    return pieces, owner


def exchange_ranges(pieces, owner, sizes, my_scores, device):
    """my_scores: {(variable, first, count): float32 tensor on `device`} for this rank's pieces.  One all_to_all_single moves
    every piece to its variable's owner.  -> {variable: float32 tensor [sizes[variable]]} for the variables this rank owns."""
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)
    send_counts = [0] * world
    for (v, first, count) in pieces[rank]:
        send_counts[owner[v]] += count
    recv_counts = [sum(c for (v, f, c) in pieces[r] if owner[v] == rank) for r in range(world)]
    # send buffer: grouped by destination, pieces in plan order
    send = torch.empty(max(1, sum(send_counts)), dtype=torch.float32, device=device)
    off = [0] * world
    base = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    for (v, first, count) in pieces[rank]:
        d = owner[v]
        send[int(base[d]) + off[d]: int(base[d]) + off[d] + count] = my_scores[(v, first, count)]
        off[d] += count
    recv = torch.empty(max(1, sum(recv_counts)), dtype=torch.float32, device=device)
    if world > 1:
        dist.all_to_all_single(recv[:sum(recv_counts)], send[:sum(send_counts)], output_split_sizes=recv_counts, input_split_sizes=send_counts)
    else:
        recv[:sum(recv_counts)] = send[:sum(send_counts)]
    out = {v: torch.empty(int(sizes[v]), dtype=torch.float32, device=device) for v in range(len(sizes)) if owner[v] == rank}
    pos = 0
    for r in range(world):
        for (v, first, count) in pieces[r]:
            if owner[v] == rank:
                out[v][first:first + count] = recv[pos:pos + count]
                pos += count
    return out


class PeerScoreBoard:
    """The exchange step of the range shards without a collective: every rank owns one device buffer holding the raw-score
    arrays of the variables it owns (``urlgpu_peer_alloc``) and maps the buffers of all other ranks (CUDA IPC, peer access
    over NVLink / NVSwitch).  ``target(v, first)`` is then the address — local or on a peer GPU — that
    ``urlgpu_score_range(..., out_device_ptr=...)`` scores a piece into, so the scoring kernels store straight into the
    owner's memory while they compute; ``fence()`` (stream sync + barrier) replaces the all-to-all of
    :func:`exchange_ranges`.  One process per GPU on one node; needs the NCCL (or any) process group only for the handle
    all-gather and the barrier."""

    def __init__(self, eng, sizes, owner):
        self.eng, self.sizes, self.owner = eng, [int(s) for s in sizes], list(owner)
        self.world, self.rank = (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)
        self.offset, fill = [0] * len(sizes), [0] * self.world
        for v, s in enumerate(self.sizes):         # every rank computes the same layout
            self.offset[v] = fill[self.owner[v]]
            fill[self.owner[v]] += (s + 63) // 64 * 64
        self.base = [None] * self.world
        ok, why, handle = 1, "", bytes(64)
        try:
            self.base[self.rank], handle = eng.peer_alloc(4 * max(1, fill[self.rank]))
        except Exception as ex:      # reported below, on every rank
            ok, why = 0, repr(ex)
        if self.world > 1:
            g = _side_group()
            mine = torch.tensor(list(handle), dtype=torch.uint8)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=g)
            for r in range(self.world):
                if r != self.rank and ok and bool(allh[r].any()):
                    try:
                        self.base[r] = eng.peer_open(bytes(allh[r].tolist()))
                    except Exception as ex:
                        ok, why = 0, repr(ex)
                elif r != self.rank:
                    ok = 0
            flag = torch.tensor([ok], dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=g)      # agreement: all proceed or all give up
            ok = int(flag.item())
        if not ok:
            self._release()
            raise PeerMemoryUnavailable("peer score board: " + (why or "another rank could not allocate or map it"))

    def target(self, v: int, first: int = 0) -> int:
        """device address of entry ``first`` of variable v's raw-score array, in its owner's memory"""
        return self.base[self.owner[v]] + 4 * (self.offset[v] + int(first))

    def fence(self):
        """all pieces have landed: this rank's kernels are done and so are everybody else's"""
        self.eng.synchronize()
        if self.world > 1:
            dist.barrier()

    def _release(self):
        for r in range(self.world):
            if self.base[r] is not None:
                try:
                    (self.eng.peer_free if r == self.rank else self.eng.peer_close)(self.base[r])
                except Exception:
                    pass
            self.base[r] = None

    def close(self):
        self.fence()
        self._release()


_BOARD = {}
_SIDE = {}


def _side_group():
    """A gloo (CPU, TCP on localhost) group for the few bytes of control traffic of the peer-memory paths: counts, handles and
    barriers.  Measured on 4 and 8 B200s: the first NCCL collective after a ~300 ms compute phase takes 55-60 ms whatever its
    size (later ones < 1 ms), which was most of the gather's cost; the gloo round trips take ~0.2 ms and involve no GPU."""
    g = _SIDE.get("g")
    if g is None:
        g = dist.new_group(backend="gloo")      # collective: every rank reaches its first peer-memory gather together
        _SIDE["g"] = g
    return g


def _cache_board(eng, total: int, words: int, dst: int):
    """rank dst's device landing area for gathered caches ([cap * words] mask words, then [cap] scores), mapped by every
    rank; (re)allocated collectively when `total` entries do not fit — every rank sees the same `total`, so all take the
    same branch.  -> (address of the area in THIS process, cap)"""
    world, rank = dist.get_world_size(), dist.get_rank()
    b = _BOARD.get(dst)
    if b is not None and b["cap"] >= total and b["words"] == words:
        return b["ptr"], b["cap"]
    g = _side_group()
    if b is not None:
        dist.barrier(group=g)
        (b["eng"].peer_free if rank == dst else b["eng"].peer_close)(b["ptr"])
        _BOARD.pop(dst, None)
    cap = int(total * 1.5) + 4096
    handle = torch.zeros(64, dtype=torch.uint8)
    ptr, ok, why = None, 1, ""
    if rank == dst:
        try:
            ptr, h = eng.peer_alloc(cap * (8 * words + 4))
            handle = torch.tensor(list(h), dtype=torch.uint8)
        except Exception as ex:
            ok, why = 0, repr(ex)
    dist.broadcast(handle, src=dst, group=g)
    if rank != dst:
        if bool(handle.any()):
            try:
                ptr = eng.peer_open(bytes(handle.tolist()))
            except Exception as ex:
                ok, why = 0, repr(ex)
        else:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=g)      # agreement: all proceed or all give up
    if int(flag.item()) == 0:
        if ptr is not None:
            try:
                (eng.peer_free if rank == dst else eng.peer_close)(ptr)
            except Exception:
                pass
        raise PeerMemoryUnavailable("gather board: " + (why or "another rank could not allocate or map it"))
    _BOARD[dst] = {"ptr": ptr, "cap": cap, "words": words, "eng": eng}
    return ptr, cap


def gather_results_p2p(eng, results: dict, p: int, words: int, dst: int = 0, owner=None, shift: int = 0, copy: bool = True):
    """:func:`gather_caches` without the host bounce: ``results`` = {variable: Result} still on the device (the variables this
    rank owns, local numbering; ``shift`` is added to every variable index, for a rank whose data set is one block of a larger
    one).  Every rank writes its compacted caches straight into rank ``dst``'s device memory over NVLink
    (``urlgpu_result_fetch_device`` into a CUDA-IPC mapping), one barrier, and ``dst`` copies the whole area to page-locked
    host memory once.  Returns {global variable: (masks uint64 [n, words], scores float32 [n])} on ``dst``, None elsewhere."""
    import os, time
    dbg = os.environ.get("URLGPU_GATHER_TIMING")
    tt = [time.perf_counter()]
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(p, dtype=torch.int64)
    owned = sorted(results)
    for v in owned:
        counts[v + shift] = results[v].count()
    tt.append(time.perf_counter())
    g = _side_group()
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts, group=g)      # CPU tensors over gloo: see _side_group
    all_counts = torch.stack(all_counts).numpy()
    per_rank = all_counts.sum(axis=1)
    base = np.concatenate([[0], np.cumsum(per_rank)]).astype(np.int64)
    total = int(per_rank.sum())
    tt.append(time.perf_counter())
    ptr, cap = _cache_board(eng, total, words, dst)
    off = int(base[rank])
    for v in owned:
        off += results[v].fetch_device(ptr + 8 * words * off, ptr + 8 * words * cap + 4 * off, words, shift)
    tt.append(time.perf_counter())
    dist.barrier(group=g)     # every rank's stores have completed (fetch_device returns after its copy stream drained)
    tt.append(time.perf_counter())
    if dbg and rank in (0, 1):
        print("[gather_results_p2p rank %d] counts %.1f ms, all_gather %.1f, stores %.1f, barrier %.1f" % (
            rank, *[1e3 * (tt[i + 1] - tt[i]) for i in range(4)]), flush=True)
    if rank != dst:
        return None
    pin = torch.cuda.is_available()
    land_m = _landing("recv_m", max(total, 1) * words, torch.int64, pin).view(-1, words)
    land_s = _landing("recv_s", max(total, 1), torch.float32, pin)
    lm, ls = land_m.numpy(), land_s.numpy()
    eng.copy_to_host(lm, ptr, 8 * words * total)
    eng.copy_to_host(ls, ptr + 8 * words * cap, 4 * total)
    if dbg:
        print("[gather_results_p2p rank %d] D2H of %d bytes %.1f ms" % (rank, (8 * words + 4) * total, 1e3 * (time.perf_counter() - tt[-1])), flush=True)
    lm = lm.view(np.uint64)
    out = {}
    for r in range(world):
        o = int(base[r])
        for v in range(p):
            k = int(all_counts[r, v])
            if k == 0 and (owner[v] if owner is not None else v % world) != r:
                continue
            out[v] = (lm[o:o + k].copy(), ls[o:o + k].copy()) if copy else (lm[o:o + k], ls[o:o + k])
            o += k
    return out


def release_boards():
    """free / unmap the gather boards (collective: call on every rank before the engines close)"""
    if not _BOARD:
        return
    dist.barrier(group=_side_group())
    rank = dist.get_rank()
    for dst, b in list(_BOARD.items()):
        (b["eng"].peer_free if rank == dst else b["eng"].peer_close)(b["ptr"])
    _BOARD.clear()
