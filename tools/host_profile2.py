"""host-side enqueue time of every variable (URLGPU_DEBUG_TIMING=1 prints the phases on stderr)"""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("urlearning-cpp_b200")
codes, card, edges, _ = pkg.datagen.discrete_bn(p=60, n=1_000_000, seed=4)
eng = pkg.Engine(0)
eng.set_discrete(codes, card)
nbs = [pkg.two_hop_neighbors(edges, 60, v) for v in range(60)]
for rep in range(4):
    if rep == 3: eng.close(); break
    eng.synchronize()
    print(f"=== pass {rep}", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    for v in range(60):
        eng.score_variable(v, nbs[v], 11, pkg.BIC, flags=pkg.PRUNE_DOMINATED).free()
    t1 = time.perf_counter()
    eng.synchronize()
    print(f"pass {rep}: enqueue {1e3*(t1-t0):.1f} ms total {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
