"""world_size-2 gloo test of the multi-process path on CPU: striping, broadcast of the input, gather of the
per-variable caches to rank 0 and the .pss rank 0 writes.  The scorer plugged in here is the CPU oracle (test
infrastructure); on the GPU box the same plumbing carries liburlgpu results (tests/test_gpu_multi.py, bench.py)."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("urlearning-cpp_b200")
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    import oracle_lib as orc
    os.chdir(os.path.join(ROOT, "tests"))
    if rank == 0:
        t = orc.Table("data/hepatitis.clean.csv", has_header=True)
        codes, card, names = torch.from_numpy(t.codes()), torch.from_numpy(t.card), t.names
        meta = [dict(p=t.p, n=t.n, names=names)]
    else:
        codes = card = None
        meta = [None]
    dist.broadcast_object_list(meta, src=0)
    p, n = meta[0]["p"], meta[0]["n"]
    codes = D.broadcast_tensor(codes, (p, n), torch.uint8, "cpu").numpy()
    card = D.broadcast_tensor(card, (p,), torch.int32, "cpu").numpy()
    K = pkg.effective_max_parents(0, p, n, True)
    costs = [D.family_cost(card, v, pkg.two_hop_neighbors(None, p, v), K) for v in range(p)]
    owner = D.assign_lpt(costs, world) if os.environ.get("URLGPU_TEST_LPT") == "1" else None
    local = {}
    for v in ([v for v in range(p) if owner[v] == rank] if owner else D.stripe(p, rank, world)):
        nb = pkg.two_hop_neighbors(None, p, v)
        masks = orc.enumerate_sets(v, nb, p, K)
        scores = orc.bic_score_many(codes, card, v, masks, threads=1)
        order = orc.canonical_order(masks)
        local[v] = (masks[order].reshape(-1, 1), scores[order])
    caches = D.gather_caches(local, p, 1, "cpu", owner=owner)
    if rank == 0:
        assert sorted(caches) == list(range(p))
        pkg.pss.write_pss(out_path, "data/hepatitis.clean.csv", n, K, "BIC", meta[0]["names"], card, caches)
    else:
        assert caches is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_pss_identical_to_single_process(tmp_path, monkeypatch):
    for k, lpt in enumerate(("0", "1")):  # the reference's striping, then cost-balanced ownership
        monkeypatch.setenv("URLGPU_TEST_LPT", lpt)
        out = str(tmp_path / f"two_rank_{lpt}.pss")
        port = 29500 + ((os.getpid() + 7 * k) % 2000)
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert hashlib.sha256(open(out, "rb").read()).hexdigest() == GOLD["hepatitis_bic"]["sha256"]


def test_stripe_matches_reference_rule():
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    for world in (1, 2, 4, 8):
        owned = [D.stripe(60, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(60))
        assert all(v % world == r for r in range(world) for v in owned[r])


def test_lpt_assignment_is_balanced_and_deterministic():
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    pkg = importlib.import_module("urlearning-cpp_b200")
    rng = np.random.default_rng(0)
    card = rng.choice([2, 3, 4], size=60)
    edges = [0] * 60
    for i in range(60):
        for j in range(max(0, i - 4), i):
            if rng.random() < 0.6:
                edges[i] |= 1 << j
                edges[j] |= 1 << i
    costs = [D.family_cost(card, v, pkg.two_hop_neighbors(edges, 60, v), 11) for v in range(60)]
    assert all(c > 0 for c in costs)
    for world in (1, 2, 4, 8):
        owner = D.assign_lpt(costs, world)
        assert owner == D.assign_lpt(list(costs), world)
        assert set(owner) <= set(range(world))
        loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(world)]
        # LPT bound: max load <= mean + largest item
        assert max(loads) <= sum(costs) / world + max(costs) + 1e-9
        stripe_loads = [sum(costs[v] for v in D.stripe(60, r, world)) for r in range(world)]
        assert max(loads) <= max(stripe_loads) + 1e-9


def test_family_cost_matches_brute_force():
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    import itertools
    card = [2, 3, 4, 2, 3]
    nb, v, K = 0b11101, 0, 2
    cand = [2, 3, 4]
    want = sum(card[v] * np.prod([card[i] for i in s]) for l in range(K + 1) for s in itertools.combinations(cand, l))
    assert D.family_cost(card, v, nb, K) == want


def test_plan_ranges_covers_every_set_once():
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 8):
        for sizes in ([10], [5, 1000000, 7, 300000], list(rng.integers(1, 500000, size=37)), [1 << 20] * 3):
            pieces, owner = D.plan_ranges(sizes, world, min_chunk=1000)
            seen = [np.zeros(int(s), dtype=np.int32) for s in sizes]
            for r in range(world):
                for (v, first, count) in pieces[r]:
                    assert count > 0 and first + count <= sizes[v]
                    seen[v][first:first + count] += 1
            assert all((x == 1).all() for x in seen)
            assert len(owner) == len(sizes) and set(owner) <= set(range(world))
            loads = [sum(c for _, _, c in pieces[r]) for r in range(world)]
            assert max(loads) - min(loads) <= 2 * 1000 * world + max(1, sum(sizes) // world // 50 + 2000)
            assert (pieces, owner) == D.plan_ranges(list(sizes), world, min_chunk=1000)


def _range_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    sizes = [1000, 50000, 3, 20000, 12345]
    pieces, owner = D.plan_ranges(sizes, world, min_chunk=100)
    # "score" of set i of variable v = v + i / 2^20 (exact in float32 for these sizes)
    mine = {(v, f, c): (torch.arange(f, f + c, dtype=torch.float32) / 1048576.0 + v) for (v, f, c) in pieces[rank]}
    got = D.exchange_ranges(pieces, owner, sizes, mine, "cpu")
    ok = all(torch.equal(t, torch.arange(0, sizes[v], dtype=torch.float32) / 1048576.0 + v) for v, t in got.items())
    owned = sorted(got)
    res = [None] * world
    dist.all_gather_object(res, (ok, owned))
    if rank == 0:
        json.dump({"ok": all(r[0] for r in res), "owned": sorted(sum((r[1] for r in res), []))}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_range_exchange(tmp_path):
    out = str(tmp_path / "ranges.json")
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_range_worker, args=(2, port, out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r["ok"] and r["owned"] == [0, 1, 2, 3, 4]


def test_peer_score_board_layout_single_process():
    """PeerScoreBoard's address arithmetic (no process group, a stand-in engine): every variable's array starts 64 scores
    aligned inside its owner's buffer, pieces land at first * 4 bytes, and close() releases what was allocated"""
    import importlib
    D = importlib.import_module("urlearning-cpp_b200.distributed")

    class FakeEngine:
        def __init__(self):
            self.allocated, self.freed, self.syncs = [], [], 0

        def peer_alloc(self, nbytes):
            self.allocated.append(nbytes)
            return 1 << 20, bytes(64)

        def peer_free(self, ptr):
            self.freed.append(ptr)

        def synchronize(self):
            self.syncs += 1

    eng = FakeEngine()
    sizes = [100, 64, 1, 1000]
    board = D.PeerScoreBoard(eng, sizes, owner=[0, 0, 0, 0])
    assert eng.allocated == [4 * (128 + 64 + 64 + 1024)]
    base = 1 << 20
    assert [board.target(v) for v in range(4)] == [base, base + 4 * 128, base + 4 * 192, base + 4 * 256]
    assert board.target(3, 10) == base + 4 * (256 + 10)
    board.fence()
    board.close()
    assert eng.freed == [base] and eng.syncs >= 2


# ---- the control flow of the peer-memory gather on CPU: stand-in engines whose "device memory" is a file under /dev/shm ----------
class _ShmEngine:
    """peer_alloc creates a file, the 64-byte handle carries its name, peer_open maps it by name: both ranks then address the
    same bytes, as CUDA IPC does for device memory.  `fail_open` makes the mapping fail, as a box without peer access would."""
    def __init__(self, tag, fail_open=False):
        self.tag, self.fail_open, self.maps, self.next = tag, fail_open, {}, 1 << 30

    def _map(self, name, nbytes=None):
        path = "/dev/shm/" + name
        if nbytes is not None:
            with open(path, "wb") as f:
                f.truncate(nbytes)
        mm = np.memmap(path, dtype=np.uint8, mode="r+")
        base, self.next = self.next, self.next + ((len(mm) + 4095) // 4096 + 1) * 4096
        self.maps[base] = (mm, path)
        return base

    def peer_alloc(self, nbytes):
        name = "urlgpu_test_%s_%d" % (self.tag, os.getpid())
        return self._map(name, nbytes), name.encode().ljust(64, b"\0")

    def peer_open(self, handle):
        if self.fail_open:
            raise RuntimeError("peer access is not available")
        return self._map(handle.rstrip(b"\0").decode())

    def peer_close(self, ptr):
        self.maps.pop(ptr)

    def peer_free(self, ptr):
        _, path = self.maps.pop(ptr)
        os.unlink(path)

    def view(self, ptr, nbytes):
        for base, (mm, _) in self.maps.items():
            if base <= ptr < base + len(mm):
                return mm[ptr - base:ptr - base + nbytes]
        raise KeyError(ptr)

    def copy_to_host(self, host, ptr, nbytes):
        if nbytes:
            host.reshape(-1).view(np.uint8)[:nbytes] = self.view(ptr, nbytes)

    def synchronize(self):
        pass


class _FakeResult:
    def __init__(self, eng, masks, scores):
        self.eng, self.masks, self.scores = eng, np.ascontiguousarray(masks, dtype=np.uint64), np.ascontiguousarray(scores, dtype=np.float32)

    def count(self):
        return len(self.scores)

    def fetch_device(self, masks_ptr, scores_ptr, words_out=None, shift=0):
        n = len(self.scores)
        wide = np.zeros((n, words_out), dtype=np.uint64)
        wide[:, 0] = self.masks.reshape(n, -1)[:, 0] << np.uint64(shift)     # the tests shift by less than 64 - p
        self.eng.view(masks_ptr, 8 * words_out * n)[:] = wide.reshape(-1).view(np.uint8)
        self.eng.view(scores_ptr, 4 * n)[:] = self.scores.view(np.uint8)
        return n


def _p2p_worker(rank, world, port, fail):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    eng = _ShmEngine("r%d" % rank, fail_open=(fail and rank == 1))
    rng = np.random.default_rng(5)
    p, words = 6, 2
    owner = [0, 1, 0, 1, 1, 0]
    full = {v: (rng.integers(1, 1 << 20, size=(10 + 3 * v, 1)).astype(np.uint64), rng.standard_normal(10 + 3 * v).astype(np.float32)) for v in range(p)}
    mine = {v: _FakeResult(eng, *full[v]) for v in range(p) if owner[v] == rank}
    if fail:
        try:
            D.gather_results_p2p(eng, mine, p, words, owner=owner)
            raise AssertionError("the gather should have been refused on every rank")
        except D.PeerMemoryUnavailable:
            pass
        # ... and every rank can fall back to the host gather in step
        local = {v: (np.concatenate([full[v][0], np.zeros_like(full[v][0])], axis=1), full[v][1]) for v in mine}
        got = D.gather_caches(local, p, words, "cpu", owner=owner)
    else:
        for _ in range(2):           # the second round reuses the mapped board
            got = D.gather_results_p2p(eng, mine, p, words, owner=owner)
        big = {v: _FakeResult(eng, np.tile(full[v][0], (40, 1)), np.tile(full[v][1], 40)) for v in mine}
        grown = D.gather_results_p2p(eng, big, p, words, owner=owner)     # does not fit: the board is re-created collectively
        if rank == 0:
            assert all(len(grown[v][1]) == 40 * len(full[v][1]) for v in range(p))
        D.release_boards()
    if rank == 0:
        assert sorted(got) == list(range(p))
        for v in range(p):
            assert np.array_equal(got[v][0][:, 0], full[v][0][:, 0]) and not got[v][0][:, 1].any()
            assert np.array_equal(got[v][1].view(np.uint32), full[v][1].view(np.uint32))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_memory_gather_control_flow_on_cpu():
    """gather_results_p2p with stand-in engines (shared files instead of CUDA IPC): counts and barriers over the gloo side
    group, slices by rank and variable, board reuse and collective growth, the result equal to the inputs"""
    mp.spawn(_p2p_worker, args=(2, 29871 + os.getpid() % 50, False), nprocs=2, join=True)


def test_peer_memory_failure_is_agreed_on_by_all_ranks():
    """a rank that cannot map the board makes EVERY rank raise PeerMemoryUnavailable (no deadlock), and the NCCL/gloo gather
    then works in step"""
    mp.spawn(_p2p_worker, args=(2, 29931 + os.getpid() % 50, True), nprocs=2, join=True)


def _board_worker(rank, world, port, fail):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    eng = _ShmEngine("b%d" % rank, fail_open=(fail and rank == 0))
    sizes = [1000, 37, 5000, 64]
    pieces, owner = D.plan_ranges(sizes, world, min_chunk=1)
    if fail:
        try:
            D.PeerScoreBoard(eng, sizes, owner)
            raise AssertionError("the board should have been refused on every rank")
        except D.PeerMemoryUnavailable:
            pass
    else:
        board = D.PeerScoreBoard(eng, sizes, owner)
        truth = {v: np.arange(sizes[v], dtype=np.float32) + 1000 * v for v in range(len(sizes))}
        for (v, first, count) in pieces[rank]:       # "score" my pieces straight into the owners' memory
            eng.view(board.target(v, first), 4 * count)[:] = truth[v][first:first + count].view(np.uint8)
        board.fence()
        for v in range(len(sizes)):
            if owner[v] == rank:
                got = np.frombuffer(bytes(eng.view(board.target(v), 4 * sizes[v])), dtype=np.float32)
                assert np.array_equal(got, truth[v]), v
        board.close()
    dist.barrier()
    dist.destroy_process_group()


def test_peer_score_board_two_ranks_on_cpu():
    """PeerScoreBoard over two gloo ranks with the shared-file stand-in: every piece written through target() lands in its
    owner's array; a rank that cannot map makes both ranks raise PeerMemoryUnavailable"""
    mp.spawn(_board_worker, args=(2, 29991 + os.getpid() % 50, False), nprocs=2, join=True)
    mp.spawn(_board_worker, args=(2, 30051 + os.getpid() % 50, True), nprocs=2, join=True)
