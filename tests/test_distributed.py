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
