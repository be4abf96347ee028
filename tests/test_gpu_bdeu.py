"""GPU parity: BDeu (bdeu_scoring_function.cpp) through the C ABI vs the CPU oracle's restatement of the same contract
(FP64 lgamma brackets on the 2^-30 grid, exact sum, one rounding).  The device evaluates lgamma with CUDA's libm, the oracle
with glibc's: a bracket may land on the neighbouring grid point, so the comparison allows one float32 ulp on a small share
of the scores; the stored lists must be identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ulps(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def _check_variable(pkg, orc, eng, codes, card, edges, v, K, ess, flags=0):
    p = codes.shape[0]
    nb = pkg.two_hop_neighbors(edges, p, v)
    res = eng.score_variable(v, nb, K, pkg.BDEU, lam=ess, flags=flags)
    masks, scores = res.fetch()
    om = orc.enumerate_sets(v, nb, p, K)
    osc = orc.bdeu_score_many(codes, card, v, om, ess=ess)
    stored = np.array([(s < 1) if m == 0 else (s < 0) for m, s in zip(om, osc)])
    om, osc = om[stored], osc[stored]
    assert res.scored() == len(stored)
    res.free()
    if flags & pkg.PRUNE_DOMINATED:
        # the prune compares float scores: apply the oracle's prune to the DEVICE's scores of the stored sets
        full = eng.score_variable(v, nb, K, pkg.BDEU, lam=ess)
        fm, fs = full.fetch()
        full.free()
        keep = orc.prune(np.array([int(m[0]) for m in fm], dtype=np.uint64), fs, K)
        want_m = [int(m[0]) for m in fm[keep]]
        assert [int(m[0]) for m in masks] == want_m
        assert np.array_equal(scores.view(np.uint32), fs[keep].view(np.uint32))
        return len(want_m)
    order = orc.canonical_order(om)
    assert [int(m[0]) for m in masks] == [int(om[i]) for i in order]
    d = _ulps(scores, osc[order])
    assert d.max() <= 1 and (d > 0).mean() < 0.02, (int(d.max()), float((d > 0).mean()))
    return len(order)


@pytest.mark.parametrize("ess", [1.0, 10.0])
def test_hepatitis_all_variables(pkg, orc, bic_engine, data_dir, ess):
    t = orc.Table(os.path.join(data_dir, "hepatitis.clean.csv"), has_header=True)
    codes = t.codes()
    bic_engine.set_discrete(codes, t.card)
    for v in range(t.p):
        _check_variable(pkg, orc, bic_engine, codes, t.card, None, v, 3, ess)


def test_mixed_arities_tiers_and_prune(pkg, orc, bic_engine):
    """tables from a few cells (shared-memory tier) to 4^8 * 4 cells (global tier, split over configuration chunks)"""
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=10, n=20011, seed=5, arities=(2, 3, 4), window=9, max_indegree=3)
    bic_engine.set_discrete(codes, card)
    for v in (0, 4, 9):
        _check_variable(pkg, orc, bic_engine, codes, card, None, v, 8, 1.0)
        _check_variable(pkg, orc, bic_engine, codes, card, None, v, 5, 2.5, flags=pkg.PRUNE_DOMINATED)


def test_score_one_ranges_and_interleaving(pkg, orc, engine):
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=12, n=10007, seed=11, arities=(2, 3, 4))
    engine.set_discrete(codes, card)
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = int(rng.integers(12))
        k = int(rng.integers(0, 7))
        others = [i for i in range(12) if i != v]
        parents = sum(1 << int(i) for i in rng.choice(others, size=k, replace=False))
        s, _ = engine.score_one(v, parents, pkg.BDEU, lam=1.0)
        want = orc.bdeu_score_many(codes, card, v, [parents], ess=1.0)[0]
        assert _ulps(s, want).max() <= 1
        b, _ = engine.score_one(v, parents, pkg.BIC)              # the next call selects BIC again
        assert b.view(np.uint32) == orc.bic_score(codes, card, v, parents)[0].view(np.uint32)
    v, nb, K = 5, (1 << 12) - 1, 4
    total = engine.family_size(v, nb, K, pkg.BDEU)
    om = orc.enumerate_sets(v, nb, 12, K)
    want = orc.bdeu_score_many(codes, card, v, om, ess=4.0)[orc.canonical_order(om)]
    got = np.concatenate([engine.score_range(v, nb, K, pkg.BDEU, a, min(500, total - a), lam=4.0) for a in range(0, total, 500)])
    assert _ulps(got, want).max() <= 1
    # results do not depend on the path: whole family vs ranges, bit for bit
    res = engine.score_variable(v, nb, K, pkg.BDEU, lam=4.0)
    m, s = res.fetch()
    res.free()
    idx = {int(mm): i for i, mm in enumerate(om[orc.canonical_order(om)])}
    assert np.array_equal(s.view(np.uint32), got[[idx[int(mm[0])] for mm in m]].view(np.uint32))


def test_errors(pkg, engine):
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=6, n=500, seed=1)
    engine.set_discrete(codes, card)
    with pytest.raises(pkg.UrlGpuError, match="equivalent sample size"):
        engine.score_variable(0, 0b111111, 2, pkg.BDEU, lam=0.0)


def test_score_binary_bdeu(pkg, orc, data_dir, tmp_path):
    """`score -f BDeu -e 2 -p 3`: same stored sets as the oracle's file, scores within one float32 ulp of it"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "urlearning-cpp_b200", "score")
    inp = os.path.join(data_dir, "hepatitis.clean.csv")
    out, ref = str(tmp_path / "gpu.pss"), str(tmp_path / "ref.pss")
    subprocess.check_call([exe, inp, out, "-s", "-f", "BDeu", "-e", "2", "-p", "3", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "BDeu", has_header=True, max_parents=3, ess=2.0)
    gm, gv = orc.parse_pss(out)
    rm, rv = orc.parse_pss(ref)
    assert gm["score_type"] == "bdeu" and gm["ess"] == rm["ess"] == "2" and gm["parent_limit"] == "3"
    assert len(gv) == len(rv) == 20
    for (an, aa, ae), (bn, ba, be) in zip(gv, rv):
        assert an == bn and aa == ba and [e[1] for e in ae] == [e[1] for e in be]
        sa = np.array([float(e[0]) for e in ae])
        sb = np.array([float(e[0]) for e in be])
        assert np.all(np.abs(sa - sb) <= 2e-6 * np.abs(sb) + 2e-6)    # %f prints six decimals
