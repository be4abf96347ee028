#!/usr/bin/env python
"""Throughput of the other discrete scores on the configs[3] shape, one B200:
  fNML  — configs[3] itself (p=60, n=1e6, 2-hop skeleton, -p 11): the cube path with the regret table as the per-configuration table
  BIC   — the same pass, for comparison
  BDeu  — the same network with n=2e4 records and -p 6 (BDeu runs on the direct-counting kernels: one table per set, rows re-read
          per set, so it is measured at the sample sizes BDeu is used with)
Device-resident timing (wall clock around EnginePool.run with a synchronize on both sides), e2e = with every cache fetched.
Prints one JSON line per score.  Usage: python tools/measure_scores.py [--steps 3]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bench
    pkg = importlib.import_module("urlearning-cpp_b200")
    D = importlib.import_module("urlearning-cpp_b200.distributed")
    torch.cuda.set_device(0)
    wl = bench.make_bic_workload(pkg, 1)
    small_codes, small_card, _, _ = pkg.datagen.discrete_bn(p=60, n=20_000, seed=4)
    pool = pkg.EnginePool(0, 4)

    def run(name, codes, card, K, stype, lam, note):
        pool.set_discrete(np.ascontiguousarray(codes), card)
        items = [(v, wl["nbs"][v]) for v in range(60)]
        costs = [D.family_cost(card, v, wl["nbs"][v], K) for v in range(60)]
        sets = sum(bench.family_size(bin(wl["nbs"][v] & ~(1 << v)).count("1"), K) for v in range(60))
        pool.run(items, K, stype, lam=lam, flags=pkg.PRUNE_DOMINATED, costs=costs)          # warm-up
        pool.reset_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.run(items, K, stype, lam=lam, flags=pkg.PRUNE_DOMINATED, costs=costs)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        launches = pool.stats()["launches_total"] // args.steps
        t0 = time.perf_counter()
        stored = sum(pool.run(items, K, stype, lam=lam, flags=pkg.PRUNE_DOMINATED, costs=costs, fetch="pinned").values())
        torch.cuda.synchronize()
        ems = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"score": name, "workload": note, "sets_per_step": sets, "ms_per_step": ms, "value": sets / ms * 1e3, "unit": "sets/s",
                          "e2e_ms_per_step": ems, "e2e_value": sets / ems * 1e3, "stored_after_prune": int(stored), "gpu_launches": int(launches)}))

    K3 = wl["K"]
    run("BIC", wl["codes"], wl["card"], K3, pkg.BIC, 0.0, "configs[3]: p=60 n=1e6, -p 12 -> 11, prune")
    run("fNML", wl["codes"], wl["card"], K3, pkg.FNML, 0.0, "configs[3] data and skeleton, -p 11 (fNML has no log-bound), prune")
    run("BDeu", small_codes, small_card, 6, pkg.BDEU, 1.0, "configs[3] network sampled at n=2e4, -p 6, ess=1, prune (direct-counting kernels)")
    run("BIC-n2e4", small_codes, small_card, 6, pkg.BIC, 0.0, "the BDeu workload scored with BIC (cube path), for comparison")
    pool.close()


if __name__ == "__main__":
    main()
