"""Profiling aid (run under `ncu --profile-from-start off`): one cBIC variable of configs[2] (c = 29: level-A stages, K3 level B,
K4 and K5 segment DPs) and one Gram of the configs[4] shape (p = 200, n = 2e6) between cudaProfilerStart/Stop."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("urlearning-cpp_b200")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
eng = pkg.Engine(0)
if what in ("all", "k3"):
    x, _ = pkg.datagen.linear_gaussian_sem(p=30, n=100_000, seed=3)
    eng.set_continuous(x)
    nb = (1 << 30) - 1
    for _ in range(2):
        eng.score_variable(3, nb, 29, pkg.CBIC, lam=2.0, flags=pkg.PRUNE_DOMINATED).free()
    eng.synchronize()
    torch.cuda.profiler.start()
    eng.score_variable(4, nb, 29, pkg.CBIC, lam=2.0, flags=pkg.PRUNE_DOMINATED).free()
    eng.synchronize()
    torch.cuda.profiler.stop()
if what in ("all", "gram"):
    p, n = 200, 2_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    xd = torch.randn((p, n), dtype=torch.float64, device="cuda", generator=g)
    def gram():
        eng.shard_begin(xd.data_ptr(), n, p)
        s1, _ = eng.shard_moments(None)
        mean = s1 / n
        a1, a2 = eng.shard_moments(mean)
        eng.shard_finish(mean, np.sqrt((a2 - a1 * a1 / n) / (n - 1.0)), n)
    gram()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    gram()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
eng.close()
print("done")
