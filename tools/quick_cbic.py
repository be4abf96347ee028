"""quick timing of the cBIC path on config-3 data (development aid, not the bench)."""
import importlib, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("urlearning-cpp_b200")
p = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
vars_ = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 7]
x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=3)
eng = pkg.Engine(0)
eng.set_continuous(x)
eng.enable_timing(True)
for v in vars_:
    eng.reset_stats()
    t0 = time.time()
    res = eng.score_variable(v, (1 << p) - 1, p - 1, pkg.CBIC, lam=2.0, flags=pkg.PRUNE_DOMINATED)
    eng.synchronize()
    dt = time.time() - t0
    st = eng.stats()
    print(f"v={v} sets={res.scored()} wall={dt*1e3:.1f} ms cbic_ms={st['ms_cbic']:.1f} accept_ms={st['ms_accept']:.1f} prune_ms={st['ms_prune']:.1f} stored={res.count()} sets/s={res.scored()/dt:.3e}")
    res.free()
