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
mine = torch.from_numpy(codes[rank * half:(rank + 1) * half]).cuda()      # this rank's block of columns
full = torch.empty((p, n), dtype=torch.uint8, device="cuda")
dist.all_gather_into_tensor(full, mine)
torch.cuda.synchronize()
eng = pkg.Engine(local)
eng.set_discrete_device(full.data_ptr(), n, p, card)
nbs = [pkg.two_hop_neighbors(edges, p, v) for v in range(p)]
owner = D.assign_lpt([D.family_cost(card, v, nbs[v], K) for v in range(p)], world)
local_caches = {}
for v in range(p):
    if owner[v] == rank:
        res = eng.score_variable(v, nbs[v], K, pkg.BIC, flags=pkg.PRUNE_DOMINATED)
        local_caches[v] = res.fetch()
        res.free()
caches = D.gather_caches(local_caches, p, 1, "cuda", owner=owner)
if rank == 0:
    pkg.pss.write_pss(out, "synthetic.csv", n, K, "BIC", [f"V{i}" for i in range(p)], card, caches)
dist.barrier()
dist.destroy_process_group()
eng.close()
'''


def test_two_rank_nccl_pss_identical_to_single_process(pkg, engine, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
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
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", str(port), str(worker), ROOT, multi], timeout=600)
    assert hashlib.sha256(open(multi, "rb").read()).hexdigest() == hashlib.sha256(open(single, "rb").read()).hexdigest()
