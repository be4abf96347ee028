"""per-variable host-call time vs GPU time of the BIC path on config 4 (development aid)."""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("urlearning-cpp_b200")
codes, card, edges, _ = pkg.datagen.discrete_bn(p=60, n=1_000_000, seed=4)
eng = pkg.Engine(0)
eng.set_discrete(codes, card)
K = 11
nbs = [pkg.two_hop_neighbors(edges, 60, v) for v in range(60)]
def run(tag, flags=pkg.PRUNE_DOMINATED, fetch=False, verbose=False):
    eng.synchronize()
    t0 = time.perf_counter()
    host = []
    for v in range(60):
        a = time.perf_counter()
        res = eng.score_variable(v, nbs[v], K, pkg.BIC, flags=flags)
        b = time.perf_counter()
        if fetch: res.fetch()
        c = time.perf_counter()
        res.free()
        host.append((b - a, c - b))
    t1 = time.perf_counter()
    eng.synchronize()
    t2 = time.perf_counter()
    print(tag, f"enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms, sum host score {1e3*sum(h[0] for h in host):.1f} fetch {1e3*sum(h[1] for h in host):.1f}")
    if verbose:
        print(" ".join(f"{v}:{1e3*h[0]:.1f}/{1e3*h[1]:.1f}" for v, h in enumerate(host)))
run("warm1"); run("warm2")
run("async", verbose=True)
run("fetch", fetch=True, verbose=True)
eng.enable_timing(True); eng.reset_stats(); run("timed"); print(eng.stats())
