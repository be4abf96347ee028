"""where the end-to-end step spends its time: upload, pipelined score+fetch (development aid)"""
import importlib, sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("urlearning-cpp_b200")
codes, card, edges, _ = pkg.datagen.discrete_bn(p=60, n=1_000_000, seed=4)
pinned = torch.from_numpy(codes).pin_memory()
host = pinned.numpy()
eng = pkg.Engine(0)
nbs = [pkg.two_hop_neighbors(edges, 60, v) for v in range(60)]
for rep in range(5):
    eng.synchronize()
    t0 = time.perf_counter()
    eng.set_discrete(host, card)
    t1 = time.perf_counter()
    ts = tf = 0.0
    prev = None
    stored = 0
    for v in range(60):
        a = time.perf_counter()
        res = eng.score_variable(v, nbs[v], 11, pkg.BIC, flags=pkg.PRUNE_DOMINATED)
        res.prefetch()
        b = time.perf_counter()
        if prev is not None:
            stored += len(prev.fetch()[1]); prev.free()
        c = time.perf_counter()
        ts += b - a; tf += c - b
        prev = res
    stored += len(prev.fetch()[1]); prev.free()
    eng.synchronize()
    t2 = time.perf_counter()
    print(f"pass {rep}: set_discrete {1e3*(t1-t0):.1f} ms, score calls {1e3*ts:.1f} ms, fetch calls {1e3*tf:.1f} ms, total {1e3*(t2-t0):.1f} ms, stored {stored}", flush=True)
eng.close()
