"""GPU parity: fNML (fnml_scoring_function.cpp) through the C ABI vs the CPU oracle's exact-integer restatement —
bit-exact scores and stored lists on every K1 strategy and both cache layouts, single sets, range shards."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_variable(pkg, orc, eng, codes, card, edges, v, K, flags=0):
    p = codes.shape[0]
    nb = pkg.two_hop_neighbors(edges, p, v)
    res = eng.score_variable(v, nb, K, pkg.FNML, flags=flags)
    masks, scores = res.fetch()
    om = orc.enumerate_sets(v, nb, p, K)
    osc = orc.fnml_score_many(codes, card, v, om)
    stored = np.array([(s < 1) if m == 0 else (s < 0) for m, s in zip(om, osc)])   # score_calculator.cpp:57-61,111-113
    om, osc = om[stored], osc[stored]
    if flags & pkg.PRUNE_DOMINATED:
        keep = orc.prune(om, osc, K)
        om, osc = om[keep], osc[keep]
    order = orc.canonical_order(om)
    assert res.scored() == len(stored)
    assert [int(m[0]) for m in masks] == [int(om[i]) for i in order]
    assert np.array_equal(scores.view(np.uint32), osc[order].view(np.uint32))
    res.free()
    return len(order)


def test_hepatitis_all_variables(pkg, orc, bic_engine, data_dir):
    t = orc.Table(os.path.join(data_dir, "hepatitis.clean.csv"), has_header=True)
    codes = t.codes()
    bic_engine.set_discrete(codes, t.card)
    for v in range(t.p):
        _check_variable(pkg, orc, bic_engine, codes, t.card, None, v, 3)   # fNML has no log-bound on the parent limit: -p 3


def test_mixed_arities_all_strategies(pkg, orc, bic_engine):
    """children of arity 2, 3 and 4 (one regret row each), tables up to 4^8 cells, singleton configurations"""
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=11, n=20011, seed=5, arities=(2, 3, 4), window=10, max_indegree=3)
    bic_engine.set_discrete(codes, card)
    for v in (0, 4, 7, 10):
        _check_variable(pkg, orc, bic_engine, codes, card, None, v, 8)
        _check_variable(pkg, orc, bic_engine, codes, card, None, v, 8, flags=pkg.PRUNE_DOMINATED)


def test_interleaved_with_bic_and_other_arities(pkg, orc, engine):
    """the per-configuration table is selected per call: BIC and fNML calls for children of different arity may alternate,
    also with the on-device compaction of an earlier variable still pending"""
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=10, n=5000, seed=3, arities=(2, 3, 4))
    engine.set_discrete(codes, card)
    nb = (1 << 10) - 1
    pend = [(v, st, engine.score_variable(v, nb, 5, st).prefetch()) for v in range(6) for st in (pkg.FNML, pkg.BIC)]
    for v, st, res in pend:
        masks, scores = res.fetch()
        om = orc.enumerate_sets(v, nb, 10, 5)
        osc = orc.fnml_score_many(codes, card, v, om) if st == pkg.FNML else orc.bic_score_many(codes, card, v, om)
        stored = np.array([(s < 1) if m == 0 else (s < 0) for m, s in zip(om, osc)])
        order = orc.canonical_order(om[stored])
        assert np.array_equal(scores.view(np.uint32), osc[stored][order].view(np.uint32)), (v, st)
        res.free()


def test_score_one_and_ranges(pkg, orc, engine):
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=12, n=10007, seed=11, arities=(2, 3, 4))
    engine.set_discrete(codes, card)
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = int(rng.integers(12))
        k = int(rng.integers(0, 7))
        others = [i for i in range(12) if i != v]
        parents = sum(1 << int(i) for i in rng.choice(others, size=k, replace=False))
        s, _ = engine.score_one(v, parents, pkg.FNML)
        want = orc.fnml_score_many(codes, card, v, [parents])[0]
        assert s.view(np.uint32) == want.view(np.uint32)
    v, nb, K = 5, (1 << 12) - 1, 4
    total = engine.family_size(v, nb, K, pkg.FNML)
    om = orc.enumerate_sets(v, nb, 12, K)
    assert total == len(om)
    want = orc.fnml_score_many(codes, card, v, om)[orc.canonical_order(om)]
    got = np.concatenate([engine.score_range(v, nb, K, pkg.FNML, a, min(300, total - a)) for a in range(0, total, 300)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_large_n_uses_the_regret_approximation(pkg, orc, engine):
    """N > 1000: Szpankowski's approximation rows; 2e5 records"""
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=8, n=200000, seed=2, arities=(2, 3))
    engine.set_discrete(codes, card)
    for v in (0, 7):
        _check_variable(pkg, orc, engine, codes, card, None, v, 4)


def test_regret_overflow_is_refused(pkg, engine):
    """C(N, r) overflows float32 for a wide child and many records: the reference would store -inf; the engine refuses"""
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 60, size=(3, 200000)).astype(np.uint8)
    engine.set_discrete(codes, [60, 60, 60])
    with pytest.raises(pkg.UrlGpuError, match="overflows float32"):
        engine.score_variable(0, 0b111, 1, pkg.FNML)


def test_score_binary_fnml_matches_oracle_pss(pkg, orc, data_dir, tmp_path):
    """`score -f fNML -p 3` vs the oracle's restatement of score_main.cpp: identical bytes (also pruned, two workers)"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "urlearning-cpp_b200", "score")
    inp = os.path.join(data_dir, "hepatitis.clean.csv")
    out, ref = str(tmp_path / "gpu.pss"), str(tmp_path / "ref.pss")
    subprocess.check_call([exe, inp, out, "-s", "-f", "fNML", "-p", "3", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "fNML", has_header=True, max_parents=3)
    assert open(out, "rb").read() == open(ref, "rb").read()
    meta, variables = orc.parse_pss(out)
    assert meta["score_type"] == "fnml" and meta["parent_limit"] == "3" and len(variables) == 20
    subprocess.check_call([exe, inp, out, "-s", "-f", "fnml", "-p", "4", "--prune", "-t", "2", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "fNML", has_header=True, max_parents=4, prune=True)
    assert open(out, "rb").read() == open(ref, "rb").read()
