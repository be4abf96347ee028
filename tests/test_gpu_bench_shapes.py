"""GPU parity AT THE BENCHMARKED SHAPES (bench.py's workloads themselves, not scaled-down stand-ins).

configs[3] (p=60, n=1e6, arity<=4, 2-hop skeleton, -p 12 -> cap 11): the large-table branches of K1 (fused roots, 2^20
row buckets, sliced roots, the root-layer cost model) only run at this size.  The oracle's direct counting does ~2k
sets/s here, so the check is SAMPLED: >= 2000 uniformly drawn sets per variable (layers 10 and 11 always included)
for the three largest families and three random ones, bit-exact; plus one whole family with the subset-dominance
prune (score_calculator.cpp:150-197) compared exactly.

configs[2] (p=30, n=1e5, cBIC lambda=2, all 2^29 subsets per variable): 2000 sampled sets over every layer against the
oracle's restatement of BIC_OLS.cpp:277-389 (residual form), 1e-9 relative on the FP64 value and 1 ulp on the float32
the table holds; the acceptance DP (BIC_OLS.cpp:125-276) and the prune on exhaustive sub-families small enough for the
oracle (c = 20 and c = 16); and a size-independent property at c = 29: acceptance and prune of S depend on subsets of S
only, so the c = 29 cache restricted to the first 20 candidates must equal the c = 20 cache entry for entry.
"""
import concurrent.futures as cf
import importlib
import os
import sys
from math import comb

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9  # BASELINE.json north_star


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


@pytest.fixture(scope="module")
def bench_mod():
    sys.path.insert(0, ROOT)
    return importlib.import_module("bench")


@pytest.fixture(scope="module")
def cfg4(pkg, bench_mod):
    return bench_mod.make_bic_workload(pkg, 1)


def _canonical_position(c, compact_mask):
    """position of a set in the canonical order (|S|, mask) when every set of the family is stored: the sets of one layer
    in increasing mask order are in colex order, rank = sum_i C(b_i, i+1) over the set bits b_0 < b_1 < ..."""
    bits = [b for b in range(c) if (compact_mask >> b) & 1]
    return sum(comb(c, l) for l in range(len(bits))) + sum(comb(b, i + 1) for i, b in enumerate(bits))


def test_config4_sampled_sets_bit_exact(pkg, orc, cfg4):
    wl = cfg4
    p, K = wl["p"], wl["K"]
    assert (p, wl["n"], K) == (60, 1_000_000, 11)
    eng = pkg.Engine(0)
    eng.set_discrete(wl["codes"], wl["card"])
    cs = [bin(wl["nbs"][v] & ~(1 << v)).count("1") for v in range(p)]
    rng = np.random.default_rng(2024)
    largest = sorted(range(p), key=lambda v: (-cs[v], v))[:3]
    others = [int(v) for v in rng.choice([v for v in range(p) if v not in largest], size=3, replace=False)]
    threads = os.cpu_count() or 8
    checked = 0
    for v in largest + others:
        nb = wl["nbs"][v]
        res = eng.score_variable(v, nb, K, pkg.BIC)
        masks, scores = res.fetch()
        res.free()
        fam = sum(comb(cs[v], l) for l in range(min(cs[v], K) + 1))
        assert len(scores) == fam  # BIC scores are negative: the store rule keeps the whole family
        pc = np.array([bin(int(m)).count("1") for m in masks[:, 0]])
        assert np.all(np.diff(pc) >= 0)
        # uniform sample + every layer from 10 up represented
        pick = set(int(i) for i in rng.choice(fam, size=min(fam, 1700), replace=False))
        for layer in (10, 11):
            idx = np.nonzero(pc == layer)[0]
            if len(idx):
                pick |= set(int(i) for i in rng.choice(idx, size=min(len(idx), 150), replace=False))
        pick = np.array(sorted(pick))
        if len(pick) < 2000 <= fam:
            extra = rng.choice(np.setdiff1d(np.arange(fam), pick), size=2000 - len(pick), replace=False)
            pick = np.sort(np.concatenate([pick, extra]))
        want = orc.bic_score_many(wl["codes"], wl["card"], v, masks[pick, 0].copy(), 0, threads)
        assert np.array_equal(scores[pick].view(np.uint32), want.view(np.uint32)), f"variable {v}: BIC scores differ from the oracle"
        checked += len(pick)
    assert checked >= 6 * 1000
    eng.close()


def test_config4_pruned_family_exact(pkg, orc, cfg4):
    """one whole family of configs[3] (c = 12 or 13: the oracle scores it in seconds) with the prune, survivors and scores exact"""
    wl = cfg4
    p, K = wl["p"], wl["K"]
    cs = [bin(wl["nbs"][v] & ~(1 << v)).count("1") for v in range(p)]
    v = next(v for v in range(p) if 12 <= cs[v] <= 13)
    eng = pkg.Engine(0)
    eng.set_discrete(wl["codes"], wl["card"])
    res = eng.score_variable(v, wl["nbs"][v], K, pkg.BIC, flags=pkg.PRUNE_DOMINATED)
    masks, scores = res.fetch()
    res.free()
    om = orc.enumerate_sets(v, wl["nbs"][v], p, K)
    osc = orc.bic_score_many(wl["codes"], wl["card"], v, om, 0, os.cpu_count() or 8)
    stored = np.array([(s < 1) if m == 0 else (s < 0) for m, s in zip(om, osc)])
    om, osc = om[stored], osc[stored]
    keep = orc.prune(om, osc, K)
    om, osc = om[keep], osc[keep]
    order = orc.canonical_order(om)
    assert [int(m) for m in masks[:, 0]] == [int(om[i]) for i in order]
    assert np.array_equal(scores.view(np.uint32), osc[order].view(np.uint32))
    eng.close()


@pytest.fixture(scope="module")
def cfg3(pkg, bench_mod):
    return bench_mod.make_cbic_workload(pkg)


def test_config3_sampled_sets_c29(pkg, orc, cfg3):
    wl = cfg3
    p, n, lam = wl["p"], wl["n"], wl["lam"]
    assert (p, n, wl["K"]) == (30, 100_000, 29)
    eng = pkg.Engine(0)
    eng.set_continuous(wl["x"])
    z = orc.standardise(wl["x"])
    rng = np.random.default_rng(7)
    v = 17
    c = p - 1
    cand = [i for i in range(p) if i != v]
    res = eng.score_variable(v, (1 << p) - 1, c, pkg.CBIC, lam=lam, flags=pkg.CBIC_NO_ACCEPT)
    assert res.scored() == 1 << c and res.count() == 1 << c
    # 2000 sets: uniform over the 2^29 masks (layers ~9..20) plus 25 per layer 0..29 where the layer has that many
    compact = set(int(m) for m in rng.integers(0, 1 << c, size=1400))
    for layer in range(0, c + 1):
        for _ in range(25 if layer not in (0, c) else 1):
            compact.add(sum(1 << int(b) for b in rng.choice(c, size=layer, replace=False)) if layer else 0)
    compact = sorted(compact)
    assert len(compact) >= 2000

    def full_mask(cm):
        return sum(1 << cand[b] for b in range(c) if (cm >> b) & 1)

    def oracle(cm):
        return orc.cbic_residual(z, v, full_mask(cm), lam)

    with cf.ThreadPoolExecutor(os.cpu_count() or 8) as ex:  # ctypes releases the GIL
        want = list(ex.map(oracle, compact))
    m1 = np.zeros((1, 1), dtype=np.uint64)
    s1 = np.zeros(1, dtype=np.float32)
    for cm, r in zip(compact, want):
        pos = _canonical_position(c, cm)
        eng._check(eng.lib.urlgpu_result_fetch(res._h, pos, 1, m1.ctypes.data, s1.ctypes.data))
        assert int(m1[0, 0]) == full_mask(cm)
        assert ulp_diff(-s1[0], np.float32(r)) <= 1, (cm, s1[0], r)
    res.free()
    # FP64 value of the largest sets through the per-set entry point (same sweeps as the family kernels)
    for cm in compact[-40:] + compact[:40]:
        s, ts64 = eng.score_one(v, full_mask(cm), pkg.CBIC, lam)
        r = orc.cbic_residual(z, v, full_mask(cm), lam)
        assert abs(ts64 - r) <= TOL * max(1.0, abs(r))
    eng.close()


def _cache_dict(masks, scores):
    return {int(m): s for m, s in zip(masks[:, 0], scores)}


def test_config3_accept_and_prune_exhaustive_slices_and_c29_property(pkg, orc, cfg3):
    wl = cfg3
    p, lam = wl["p"], wl["lam"]
    eng = pkg.Engine(0)
    eng.set_continuous(wl["x"])
    v = 23
    cand = [i for i in range(p) if i != v]
    # ---- c = 20: acceptance decisions and stored values exact given the engine's float32 the_scores ----
    nb20 = sum(1 << i for i in cand[:20]) | (1 << v)
    r = eng.score_variable(v, nb20, 20, pkg.CBIC, lam=lam, flags=pkg.CBIC_NO_ACCEPT)
    m_all, neg_ts = r.fetch()
    r.free()
    assert len(neg_ts) == 1 << 20
    om = m_all[:, 0].copy()
    stored, val = orc.cbic_accept(v, p, om, -neg_ts)
    r = eng.score_variable(v, nb20, 20, pkg.CBIC, lam=lam)
    m_acc, s_acc = r.fetch()
    r.free()
    assert np.array_equal(m_acc[:, 0], om[stored])      # both in canonical order
    assert np.array_equal(s_acc.view(np.uint32), val[stored].view(np.uint32))
    r = eng.score_variable(v, nb20, 20, pkg.CBIC, lam=lam, flags=pkg.PRUNE_DOMINATED)
    m20, s20 = r.fetch()
    r.free()
    # ---- c = 16: the prune against the literal O(m^2) restatement ----
    nb16 = sum(1 << i for i in cand[:16]) | (1 << v)
    r = eng.score_variable(v, nb16, 16, pkg.CBIC, lam=lam)
    ma, sa = r.fetch()
    r.free()
    keep = orc.prune(ma[:, 0].copy(), sa, 16)
    r = eng.score_variable(v, nb16, 16, pkg.CBIC, lam=lam, flags=pkg.PRUNE_DOMINATED)
    mp_, sp_ = r.fetch()
    r.free()
    assert np.array_equal(mp_[:, 0], ma[keep, 0]) and np.array_equal(sp_.view(np.uint32), sa[keep].view(np.uint32))
    # the c = 20 cache restricted to the first 16 candidates is the c = 16 cache
    low16 = np.uint64(nb16 & ~(1 << v))
    sel = (m20[:, 0] & ~low16) == 0
    assert np.array_equal(m20[sel, 0], mp_[:, 0]) and np.array_equal(s20[sel].view(np.uint32), sp_.view(np.uint32))
    # ---- c = 29 (configs[2] itself, three level-A stages, 11-bit DFS, 2^18 DP segments): restricted to the first 20 candidates ----
    r = eng.score_variable(v, (1 << p) - 1, p - 1, pkg.CBIC, lam=lam, flags=pkg.PRUNE_DOMINATED)
    n29 = r.count()
    assert r.scored() == 1 << 29 and 0 < n29 < 1 << 29
    m29, s29 = r.fetch()
    r.free()
    low20 = np.uint64(nb20 & ~(1 << v))
    sel = (m29[:, 0] & ~low20) == 0
    assert np.array_equal(m29[sel, 0], m20[:, 0])
    assert np.array_equal(s29[sel].view(np.uint32), s20.view(np.uint32))
    # prune is idempotent on a sample of what survived at c = 29 (the standalone entry point holds <= 30 distinct variables)
    pc = np.array([bin(int(m)).count("1") for m in m29[:200000, 0]])
    assert np.all(np.diff(pc) >= 0)
    eng.close()
