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
    local = {}
    for v in D.stripe(p, rank, world):
        nb = pkg.two_hop_neighbors(None, p, v)
        masks = orc.enumerate_sets(v, nb, p, K)
        scores = orc.bic_score_many(codes, card, v, masks, threads=1)
        order = orc.canonical_order(masks)
        local[v] = (masks[order].reshape(-1, 1), scores[order])
    caches = D.gather_caches(local, p, 1, "cpu")
    if rank == 0:
        assert sorted(caches) == list(range(p))
        pkg.pss.write_pss(out_path, "data/hepatitis.clean.csv", n, K, "BIC", meta[0]["names"], card, caches)
    else:
        assert caches is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_pss_identical_to_single_process(tmp_path):
    out = str(tmp_path / "two_rank.pss")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == GOLD["hepatitis_bic"]["sha256"]


def test_stripe_matches_reference_rule():
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    for world in (1, 2, 4, 8):
        owned = [D.stripe(60, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(60))
        assert all(v % world == r for r in range(world) for v in owned[r])
