"""Two ranks over NCCL (needs >= 2 GPUs, skipped otherwise): the input is all-gathered from the blocks each rank holds,
variables are dealt out by predicted cost, every rank scores its share through the C ABI, the caches are gathered to
rank 0 and the .pss it writes is byte-identical to the single-process one."""
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = sys.argv[1]; out = sys.argv[2]
sys.path.insert(0, ROOT)
pkg = importlib.import_module("urlearning-cpp_b200")
D = importlib.import_module("urlearning-cpp_b200.distributed")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p, n, K = 24, 40000, 5
codes, card, edges, _ = pkg.datagen.discrete_bn(p=p, n=n, seed=7, window=3, max_indegree=2)
half = p // world
mine = torch.from_numpy(np.ascontiguousarray(codes[rank * half:(rank + 1) * half])).cuda()      # this rank's block of columns
full = torch.empty((p, n), dtype=torch.uint8, device="cuda")
dist.all_gather_into_tensor(full, mine)
torch.cuda.synchronize()
eng = pkg.Engine(local)
eng.set_discrete_device(full.data_ptr(), n, p, card)
nbs = [pkg.two_hop_neighbors(edges, p, v) for v in range(p)]
owner = D.assign_lpt([D.family_cost(card, v, nbs[v], K) for v in range(p)], world)
local_caches, kept = {}, {}
for v in range(p):
    if owner[v] == rank:
        res = eng.score_variable(v, nbs[v], K, pkg.BIC, flags=pkg.PRUNE_DOMINATED)
        local_caches[v] = res.fetch()
        kept[v] = res
caches = D.gather_caches(local_caches, p, 1, "cuda", owner=owner)
# the same gather without the host bounce: every rank writes its compacted caches into rank 0's device memory over NVLink
# (twice: the second call reuses the mapped board)
for _ in range(2):
    direct = D.gather_results_p2p(eng, kept, p, 1, owner=owner)
for r in kept.values():
    r.free()
if rank == 0:
    assert sorted(direct) == sorted(caches) == list(range(p))
    for v in range(p):
        assert np.array_equal(direct[v][0], caches[v][0]) and np.array_equal(direct[v][1].view(np.uint32), caches[v][1].view(np.uint32)), v
    pkg.pss.write_pss(out, "synthetic.csv", n, K, "BIC", [f"V{i}" for i in range(p)], card, direct)
D.release_boards()
dist.barrier()
dist.destroy_process_group()
eng.close()
'''


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_nccl_pss_identical_to_single_process(pkg, engine, tmp_path, nranks):
    """the .pss is byte-identical at 1, 2, 4 and 8 ranks (p = 24 is divisible by each)"""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    p, n, K = 24, 40000, 5
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=p, n=n, seed=7, window=3, max_indegree=2)
    engine.set_discrete(codes, card)
    caches = {}
    for v in range(p):
        res = engine.score_variable(v, pkg.two_hop_neighbors(edges, p, v), K, pkg.BIC, flags=pkg.PRUNE_DOMINATED)
        caches[v] = res.fetch()
        res.free()
    single = str(tmp_path / "single.pss")
    pkg.pss.write_pss(single, "synthetic.csv", n, K, "BIC", [f"V{i}" for i in range(p)], card, caches)
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    multi = str(tmp_path / "multi.pss")
    port = 29600 + os.getpid() % 300
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
                           "--master-port", str(port + nranks), str(worker), ROOT, multi], timeout=600)
    assert hashlib.sha256(open(multi, "rb").read()).hexdigest() == hashlib.sha256(open(single, "rb").read()).hexdigest()


RANGE_WORKER = r'''
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = sys.argv[1]; out = sys.argv[2]; exchange = sys.argv[3]
sys.path.insert(0, ROOT)
pkg = importlib.import_module("urlearning-cpp_b200")
D = importlib.import_module("urlearning-cpp_b200.distributed")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p, n, K, lam = 70, 6000, 3, 2.0
x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=21)
eng = pkg.Engine(local)
# rows sharded for the Gram (config-5 protocol), Gram summed in rank order
lo, hi = rank * n // world, (rank + 1) * n // world
eng.shard_begin(x[:, lo:hi].copy())
def allsum(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    acc = parts[0].clone()
    for q in parts[1:]:
        acc += q
    return acc.cpu().numpy()
mean = allsum(eng.shard_moments(None)[0]) / n
a1, a2 = eng.shard_moments(mean)
S1, S2 = allsum(a1), allsum(a2)
eng.shard_finish(mean, np.sqrt((S2 - S1 * S1 / n) / (n - 1.0)), n)
eng.set_gram(allsum(eng.gram()), n)
nbs = [(1 << p) - 1] * p
sizes = [eng.family_size(v, nbs[v], K, pkg.CBIC) for v in range(p)]
pieces, owner = D.plan_ranges(sizes, world)
local_caches = {}
if exchange == "p2p":
    # every piece is scored straight into its owner's memory over NVLink (CUDA IPC peer mapping); no collective
    board = D.PeerScoreBoard(eng, sizes, owner)
    for (v, first, count) in pieces[rank]:
        eng.score_range(v, nbs[v], K, pkg.CBIC, first, count, lam=lam, out_device_ptr=board.target(v, first))
    board.fence()
    for v in range(p):
        if owner[v] == rank:
            res = eng.result_from_scores(v, nbs[v], K, pkg.CBIC, board.target(v), n=sizes[v], flags=pkg.PRUNE_DOMINATED)
            local_caches[v] = res.fetch()
            res.free()
    board.close()
else:
    mine = {}
    for (v, first, count) in pieces[rank]:
        t = torch.empty(count, dtype=torch.float32, device="cuda")
        eng.score_range(v, nbs[v], K, pkg.CBIC, first, count, lam=lam, out_device_ptr=t.data_ptr())
        mine[(v, first, count)] = t
    eng.synchronize()
    full = D.exchange_ranges(pieces, owner, sizes, mine, "cuda")
    for v, t in full.items():
        res = eng.result_from_scores(v, nbs[v], K, pkg.CBIC, t.data_ptr(), n=sizes[v], flags=pkg.PRUNE_DOMINATED)
        local_caches[v] = res.fetch()
        res.free()
caches = D.gather_caches(local_caches, p, 2, "cuda", owner=owner)
if rank == 0:
    pkg.pss.write_pss(out, "synthetic.csv", n, K, "cBIC", [f"V{i}" for i in range(p)], [n] * p, caches)
dist.barrier()
dist.destroy_process_group()
eng.close()
'''


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
def test_two_rank_nccl_parent_set_range_shards(pkg, tmp_path, exchange):
    """cBIC over 69 candidates per variable, sharded by (variable, parent-set range) over two ranks: row-sharded Gram, ranges
    scored into NCCL buffers and moved with one all-to-all ("nccl"), or scored straight into the owner's memory over NVLink
    through a CUDA-IPC peer mapping with only a barrier ("p2p", urlgpu_peer_*); filters, gather to rank 0: the .pss equals
    the one-process one (two ranks: the one-process run sums the same two Gram shards in the same order, so the Gram bits
    are the same)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    p, n, K, lam = 70, 6000, 3, 2.0
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=21)
    # the single-process run uses the same two-shard Gram protocol so that both see the same Gram bits
    engs = [pkg.Engine(0), pkg.Engine(0)]
    for r, e in enumerate(engs):
        e.shard_begin(x[:, r * n // 2:(r + 1) * n // 2].copy())
    mean = (engs[0].shard_moments(None)[0] + engs[1].shard_moments(None)[0]) / n
    parts = [e.shard_moments(mean) for e in engs]
    S1, S2 = parts[0][0] + parts[1][0], parts[0][1] + parts[1][1]
    for e in engs:
        e.shard_finish(mean, np.sqrt((S2 - S1 * S1 / n) / (n - 1.0)), n)
    g = engs[0].gram() + engs[1].gram()
    eng = engs[0]
    eng.set_gram(g, n)
    caches = {}
    for v in range(p):
        res = eng.score_variable(v, (1 << p) - 1, K, pkg.CBIC, lam=lam, flags=pkg.PRUNE_DOMINATED)
        caches[v] = res.fetch()
        res.free()
    single = str(tmp_path / "single.pss")
    pkg.pss.write_pss(single, "synthetic.csv", n, K, "cBIC", [f"V{i}" for i in range(p)], [n] * p, caches)
    for e in engs:
        e.close()
    worker = tmp_path / "range_worker.py"
    worker.write_text(RANGE_WORKER)
    multi = str(tmp_path / "multi.pss")
    port = 29900 + os.getpid() % 300
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port + (7 if exchange == "p2p" else 0)), str(worker), ROOT, multi, exchange], timeout=600)
    assert hashlib.sha256(open(multi, "rb").read()).hexdigest() == hashlib.sha256(open(single, "rb").read()).hexdigest()
