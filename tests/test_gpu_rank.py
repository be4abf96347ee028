"""GPU parity for candidate families with MORE THAN 30 candidates (rank-space layout of the score cache).

The reference enumerates any neighbourhood below 64 variables (score_calculator.cpp:65-120); its default invocation
without -k makes every variable a candidate, so a 40-63 column CSV under the BIC parent cap is an ordinary input there.
Round 1 of this engine refused it (dense 2^c table).  Also covers cBIC over up to 199 candidates (BASELINE configs[4]:
p=200, degree-16 skeleton, explicit -p K), which the reference's 64-bit varsets cannot represent at all.
"""
import os
import subprocess
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def _oracle_bic_cache(orc, codes, card, v, nb, p, K, prune):
    om = orc.enumerate_sets(v, nb, p, K)
    osc = orc.bic_score_many(codes, card, v, om)
    stored = np.array([(s < 1) if m == 0 else (s < 0) for m, s in zip(om, osc)])
    om, osc = om[stored], osc[stored]
    if prune:
        keep = orc.prune(om, osc, K)
        om, osc = om[keep], osc[keep]
    order = orc.canonical_order(om)
    return om[order], osc[order]


@pytest.mark.parametrize("p,n,K,arities", [(50, 500, 3, (2, 3)), (62, 300, 2, (2, 3, 4)), (36, 4001, 4, (2, 3, 18))])
def test_bic_more_than_30_candidates(pkg, orc, engine, p, n, K, arities):
    """no skeleton: every other variable is a candidate (c = p - 1 > 30).  The (36, K=4, arity 18) case has tables of
    18^5 = 1.9M cells: shared-memory tiers and the global (L2 scratch) tier all run."""
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=p, n=n, seed=p, arities=arities, window=5, max_indegree=2)
    engine.set_discrete(codes, card)
    nb = (1 << p) - 1
    for v in (0, p // 2, p - 1):
        for flags in (0, pkg.PRUNE_DOMINATED):
            res = engine.score_variable(v, nb, K, pkg.BIC, flags=flags)
            masks, scores = res.fetch()
            assert res.scored() == sum(comb(p - 1, l) for l in range(K + 1))
            res.free()
            om, osc = _oracle_bic_cache(orc, codes, card, v, nb, p, K, bool(flags))
            assert [int(m) for m in masks[:, 0]] == [int(m) for m in om]
            assert np.array_equal(scores.view(np.uint32), osc.view(np.uint32))


def test_score_binary_50_variables_no_skeleton(pkg, orc, tmp_path):
    """the default invocation of the reference's `score` on a 50-column CSV (BIC cap (int)ln(2N/ln N) = 4 at N = 300)"""
    p, n = 50, 300
    codes, card, _, _ = pkg.datagen.discrete_bn(p=p, n=n, seed=77, arities=(2, 3), window=4, max_indegree=2)
    inp = str(tmp_path / "d50.csv")
    pkg.datagen.write_csv(inp, codes, header=[f"V{i}" for i in range(p)])
    exe = os.path.join(ROOT, "urlearning-cpp_b200", "score")
    out, ref = str(tmp_path / "gpu.pss"), str(tmp_path / "ref.pss")
    subprocess.check_call([exe, inp, out, "-s", "-f", "BIC", "-p", "3", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "BIC", has_header=True, max_parents=3)
    assert open(out, "rb").read() == open(ref, "rb").read()
    subprocess.check_call([exe, inp, out, "-s", "-f", "BIC", "-p", "2", "--prune", "-t", "2", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "BIC", has_header=True, max_parents=2, prune=True)
    assert open(out, "rb").read() == open(ref, "rb").read()


@pytest.mark.parametrize("p,n,K", [(40, 2000, 3), (33, 1500, 4)])
def test_cbic_more_than_30_candidates_vs_oracle(pkg, orc, engine, p, n, K):
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=p)
    engine.set_continuous(x)
    z = orc.standardise(x)
    rng = np.random.default_rng(p)
    nb = (1 << p) - 1
    for v in (1, p - 1):
        om = orc.enumerate_sets(v, nb, p, K)
        order = orc.canonical_order(om)
        om = om[order]
        res = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.CBIC_NO_ACCEPT)
        masks, neg_ts = res.fetch()
        res.free()
        assert np.array_equal(masks[:, 0], om)
        gpu_ts = -neg_ts
        for i in rng.choice(len(om), size=400, replace=False):
            r = orc.cbic_residual(z, v, int(om[i]), 2.0)
            assert ulp_diff(gpu_ts[i], np.float32(r)) <= 1
            s, ts64 = engine.score_one(v, int(om[i]), pkg.CBIC, 2.0)
            assert abs(ts64 - r) <= TOL * max(1.0, abs(r)) and ulp_diff(-s, gpu_ts[i]) == 0
        stored, val = orc.cbic_accept(v, p, om, gpu_ts)
        res = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0)
        m2, s2 = res.fetch()
        res.free()
        assert np.array_equal(m2[:, 0], om[stored]) and np.array_equal(s2.view(np.uint32), val[stored].view(np.uint32))
        keep = orc.prune(om[stored], val[stored], K)
        res = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.PRUNE_DOMINATED)
        m3, s3 = res.fetch()
        res.free()
        assert np.array_equal(m3[:, 0], om[stored][keep]) and np.array_equal(s3.view(np.uint32), val[stored][keep].view(np.uint32))


def test_cbic_199_candidates_restriction_property_and_samples(pkg, orc):
    """BASELINE configs[4] shape: p = 200, every other variable a candidate, -p 3 (1.3e6 sets per variable).  Sampled sets
    against the oracle on the relabelled sub-problem; and, because acceptance and prune of S only look at subsets of S, the
    cache restricted to the first 24 candidates must equal the 24-candidate family's cache computed in the DENSE layout
    (which the other tests hold to the oracle)."""
    from conftest import engine_with_env
    p, n, K = 200, 3000, 3
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=5, mean_indegree=2.0)
    eng = pkg.Engine(0)
    dense = engine_with_env(pkg, {"URLGPU_LAYOUT": "dense"})
    eng.set_continuous(x)
    dense.set_gram(eng.gram(), n)
    rng = np.random.default_rng(1)
    allbits = (1 << p) - 1
    for v in (0, 137):
        cand = [i for i in range(p) if i != v]
        res = eng.score_variable(v, allbits, K, pkg.CBIC, lam=2.0, flags=pkg.CBIC_NO_ACCEPT)
        total = sum(comb(p - 1, l) for l in range(K + 1))
        assert res.scored() == total and res.count() == total
        m1 = np.zeros((1, 4), dtype=np.uint64)
        s1 = np.zeros(1, dtype=np.float32)
        for _ in range(150):
            l = int(rng.integers(0, K + 1))
            pos = sorted(int(b) for b in rng.choice(p - 1, size=l, replace=False))
            idx = sum(comb(p - 1, j) for j in range(l)) + sum(comb(b, i + 1) for i, b in enumerate(pos))
            eng._check(eng.lib.urlgpu_result_fetch(res._h, idx, 1, m1.ctypes.data, s1.ctypes.data))
            members = [cand[b] for b in pos]
            assert pkg.words_to_mask(m1[0]) == sum(1 << i for i in members)
            sub = sorted(members + [v])
            zs = orc.standardise(x[sub])
            r = orc.cbic_residual(zs, sub.index(v), sum(1 << sub.index(i) for i in members), 2.0)
            # the oracle standardises the sub-problem's columns itself: same columns, same arithmetic
            assert ulp_diff(-s1[0], np.float32(r)) <= 1
        res.free()
        for flags in (0, pkg.PRUNE_DOMINATED):
            res = eng.score_variable(v, allbits, K, pkg.CBIC, lam=2.0, flags=flags)
            mw, sw = res.fetch()
            res.free()
            first24 = sum(1 << i for i in cand[:24])
            rd = dense.score_variable(v, first24 | (1 << v), K, pkg.CBIC, lam=2.0, flags=flags)
            md, sd = rd.fetch()
            rd.free()
            inside = np.array([pkg.words_to_mask(row) & ~first24 == 0 for row in mw])
            assert np.array_equal(mw[inside], md)
            assert np.array_equal(sw[inside].view(np.uint32), sd.view(np.uint32))
            pc = [bin(pkg.words_to_mask(row)).count("1") for row in mw[:: max(1, len(mw) // 5000)]]
            assert pc == sorted(pc)
    eng.close()
    dense.close()


def test_layouts_agree_bit_for_bit(pkg):
    """same family through the dense 2^c table and through the rank-space table: identical caches (BIC and cBIC, all filters)"""
    from conftest import engine_with_env
    d = engine_with_env(pkg, {"URLGPU_LAYOUT": "dense"})
    r = engine_with_env(pkg, {"URLGPU_LAYOUT": "rank"})
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=22, n=70001, seed=3, window=5, max_indegree=3)
    x, _ = pkg.datagen.linear_gaussian_sem(p=22, n=5000, seed=9)
    for e in (d, r):
        e.set_discrete(codes, card)
        e.set_continuous(x)
    for v, K in ((0, 3), (11, 5), (21, 8)):
        nb = pkg.two_hop_neighbors(edges, 22, v) | sum(1 << i for i in range(0, 22, 3))
        for st, flags in ((pkg.BIC, 0), (pkg.BIC, pkg.PRUNE_DOMINATED), (pkg.CBIC, 0), (pkg.CBIC, pkg.PRUNE_DOMINATED), (pkg.CBIC, pkg.CBIC_NO_ACCEPT)):
            a = d.score_variable(v, nb, K, st, lam=2.0, flags=flags)
            b = r.score_variable(v, nb, K, st, lam=2.0, flags=flags)
            (ma, sa), (mb, sb) = a.fetch(), b.fetch()
            assert a.scored() == b.scored()
            a.free()
            b.free()
            assert np.array_equal(ma, mb) and np.array_equal(sa.view(np.uint32), sb.view(np.uint32))
    d.close()
    r.close()


def test_parent_set_range_shards_reassemble_to_the_same_cache(pkg):
    """(variable, parent-set range) shards: ranges of a family's canonical numbering scored by two contexts
    (two ranks' stand-ins), raw scores concatenated, filters applied by the owner: identical to urlgpu_score_variable."""
    a, b = pkg.Engine(0), pkg.Engine(0)
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=40, n=30011, seed=6, window=5, max_indegree=3)
    x, _ = pkg.datagen.linear_gaussian_sem(p=40, n=4000, seed=8)
    for e in (a, b):
        e.set_discrete(codes, card)
        e.set_continuous(x)
    allbits = (1 << 40) - 1
    cases = [(pkg.BIC, 3, allbits, 3), (pkg.BIC, 17, pkg.two_hop_neighbors(edges, 40, 17), 6), (pkg.CBIC, 9, allbits, 4),
             (pkg.CBIC, 30, sum(1 << i for i in range(0, 40, 2)), 7)]
    for st, v, nb, K in cases:
        total = a.family_size(v, nb, K, st)
        cut1, cut2 = total // 3, total // 3 + 1
        parts = [a.score_range(v, nb, K, st, 0, cut1, lam=2.0), b.score_range(v, nb, K, st, cut1, cut2 - cut1, lam=2.0),
                 a.score_range(v, nb, K, st, cut2, total - cut2, lam=2.0)]
        raw = np.concatenate(parts)
        assert len(raw) == total
        for flags in (0, pkg.PRUNE_DOMINATED):
            want = a.score_variable(v, nb, K, st, lam=2.0, flags=flags)
            got = b.result_from_scores(v, nb, K, st, raw, flags=flags)
            (mw, sw), (mg, sg) = want.fetch(), got.fetch()
            assert want.scored() == got.scored() == total
            want.free()
            got.free()
            assert np.array_equal(mw, mg) and np.array_equal(sw.view(np.uint32), sg.view(np.uint32))
    with pytest.raises(pkg.UrlGpuError, match="exceeds the family"):
        a.score_range(3, allbits, 3, pkg.BIC, 10, a.family_size(3, allbits, 3, pkg.BIC))
    a.close()
    b.close()


def test_family_parts_merge_to_the_whole(pkg):
    """urlgpu_score_part: a variable's K1 work split into sub-forests of its root tables (BIC cube path) or ranges; the parts
    are disjoint, cover the family, and their int32-MIN merge gives the cache of the one-call path, bit for bit"""
    eng = pkg.Engine(0)
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=20, n=70001, seed=12, window=7, max_indegree=3)
    x, _ = pkg.datagen.linear_gaussian_sem(p=20, n=4000, seed=2)
    eng.set_discrete(codes, card)
    eng.set_continuous(x)
    NOT_SCORED = np.uint32(0x7FC0BEEF)
    for st, v, nb, K in [(pkg.BIC, 10, pkg.two_hop_neighbors(edges, 20, 10), 9), (pkg.BIC, 3, (1 << 20) - 1, 4), (pkg.BIC, 19, pkg.two_hop_neighbors(edges, 20, 19), 11),
                         (pkg.CBIC, 7, (1 << 20) - 1, 5)]:
        total = eng.family_size(v, nb, K, st)
        for parts in (2, 3, 8):
            arrs = [eng.score_part(v, nb, K, st, part, parts, lam=2.0) for part in range(parts)]
            bits = np.stack([a.view(np.uint32) for a in arrs])
            scored = bits != NOT_SCORED
            assert np.all(scored.sum(axis=0) == 1), "every set belongs to exactly one part"
            if st == pkg.BIC and parts == 2 and total > 50000:
                assert all(s.sum() > total // 8 for s in scored), "both halves of a big family carry a real share of it"
            merged = np.min(np.stack([a.view(np.int32) for a in arrs]), axis=0).view(np.float32)
            for flags in (0, pkg.PRUNE_DOMINATED):
                want = eng.score_variable(v, nb, K, st, lam=2.0, flags=flags)
                got = eng.result_from_scores(v, nb, K, st, merged, flags=flags)
                (mw, sw), (mg, sg) = want.fetch(), got.fetch()
                want.free()
                got.free()
                assert len(sw) > 0 and np.array_equal(mw, mg) and np.array_equal(sw.view(np.uint32), sg.view(np.uint32))
    eng.close()
