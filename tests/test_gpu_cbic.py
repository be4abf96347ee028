"""GPU parity: continuous cBIC (Gram + Schur sweeps + acceptance DP + prune) vs the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-9  # BASELINE.json north_star: scores within 1e-9 relative (checked on the FP64 value before the float32 rounding)


def ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, -(a & 0x7FFFFFFF), a)
    b = np.where(b < 0, -(b & 0x7FFFFFFF), b)
    return np.abs(a - b)


def test_gram_matches_extended_precision(pkg, orc, engine):
    x, _ = pkg.datagen.linear_gaussian_sem(p=13, n=20001, seed=2)
    x[3] += 1e3  # a large mean must not hurt (two-pass centring)
    engine.set_continuous(x)
    g = engine.gram()
    ref = orc.gram(orc.standardise(x))
    assert np.allclose(np.diag(g), 20000.0, rtol=1e-12)
    assert np.max(np.abs(g - ref)) <= 1e-12 * 20001
    assert np.array_equal(g, g.T)


@pytest.mark.parametrize("fig,fn", [("Figure_1", "raw_data_8000.csv"), ("Figure_2", "raw_data_5000.csv")])
def test_figures_full_family(pkg, orc, cbic_engine, data_dir, fig, fn):
    engine = cbic_engine
    t = orc.Table(os.path.join(data_dir, fig, fn))
    x = t.values()
    assert x.shape == (4, 5000)
    engine.set_continuous(x)
    z = orc.standardise(x)
    for v in range(4):
        nb = pkg.two_hop_neighbors(None, 4, v)
        om = orc.enumerate_sets(v, nb, 4, 3)
        ts = np.array([orc.cbic_residual(z, v, int(m), 2.0) for m in om])
        for m, r in zip(om, ts):
            s, ts64 = engine.score_one(v, int(m), pkg.CBIC, 2.0)
            assert abs(ts64 - r) <= TOL * max(1.0, abs(r))
            assert ulp_diff(s, np.float32(-np.float32(r))) <= 1
        stored, val = orc.cbic_accept(v, 4, om, ts.astype(np.float32))
        res = engine.score_variable(v, nb, 3, pkg.CBIC, lam=2.0)
        masks, scores = res.fetch()
        order = [i for i in orc.canonical_order(om) if stored[i]]
        assert [int(m[0]) for m in masks] == [int(om[i]) for i in order]
        assert np.all(ulp_diff(scores, val[order]) <= 1)


@pytest.mark.parametrize("p,n,K", [(9, 2000, 8), (13, 3000, 12), (14, 1500, 4), (17, 4000, 3), (27, 3000, 3)])
def test_synthetic_exhaustive_and_limited(pkg, orc, cbic_engine, p, n, K):
    engine = cbic_engine
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=p)
    engine.set_continuous(x)
    g = engine.gram()
    z = orc.standardise(x)
    rng = np.random.default_rng(p)
    for v in (0, p // 2, p - 1):
        nb = (1 << p) - 1
        om = orc.enumerate_sets(v, nb, p, K)
        # every set's score without the acceptance filter
        res = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.CBIC_NO_ACCEPT)
        masks, neg_ts = res.fetch()
        assert res.scored() == len(om) and len(masks) == len(om)
        order = orc.canonical_order(om)
        assert [int(m[0]) for m in masks] == [int(om[i]) for i in order]
        gpu_ts = np.zeros(len(om), dtype=np.float32)
        gpu_ts[order] = -neg_ts
        # (i) against the Cholesky form on the engine's own Gram and (ii) against the reference's residual form
        sample = rng.choice(len(om), size=min(len(om), 300), replace=False)
        for i in sample:
            r1 = orc.cbic_gram(g, n, v, int(om[i]), 2.0)
            r2 = orc.cbic_residual(z, v, int(om[i]), 2.0)
            s, ts64 = engine.score_one(v, int(om[i]), pkg.CBIC, 2.0)
            assert abs(ts64 - r1) <= TOL * max(1.0, abs(r1))
            assert abs(ts64 - r2) <= TOL * max(1.0, abs(r2))
            assert ulp_diff(gpu_ts[i], np.float32(r2)) <= 1
            assert ulp_diff(-s, gpu_ts[i]) == 0
        # acceptance DP and prune: decisions bit-exact given the engine's float32 the_scores
        stored, val = orc.cbic_accept(v, p, om, gpu_ts)
        r = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0)
        m2, s2 = r.fetch()
        keep_order = [i for i in order if stored[i]]
        assert [int(m[0]) for m in m2] == [int(om[i]) for i in keep_order]
        assert np.array_equal(s2.view(np.uint32), val[keep_order].view(np.uint32))
        km, ks = om[stored], val[stored]
        keep = orc.prune(km, ks, K)
        r3 = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.PRUNE_DOMINATED)
        m3, s3 = r3.fetch()
        po = [i for i in orc.canonical_order(km) if keep[i]]
        assert [int(m[0]) for m in m3] == [int(km[i]) for i in po]
        assert np.array_equal(s3.view(np.uint32), ks[po].view(np.uint32))
        # RSS is monotone under set inclusion: the best score never gets worse... (property) prune is idempotent
        again = engine.prune(m3, s3)
        assert again.all()


def test_standalone_prune_matches_literal(pkg, orc, engine):
    rng = np.random.default_rng(3)
    for trial in range(20):
        c = int(rng.integers(3, 9))
        m = int(rng.integers(1, 1 << c))
        masks = rng.choice(1 << c, size=m, replace=False).astype(np.uint64)
        masks = (masks << np.uint64(3))  # not starting at bit 0
        scores = (-rng.integers(1, 12, size=m) * 7.25).astype(np.float32)  # many exact ties, |score| >= 4
        keep = engine.prune(masks, scores)
        ref = orc.prune(masks, scores)
        assert np.array_equal(keep, ref)


def _mec_of_best_dag(variables):
    """exhaustive order search on p=4; local score = best stored subset (sparse_parent_list.cpp:44-55)"""
    import itertools
    names = [v[0] for v in variables]
    table = [{frozenset(names.index(q) for q in pa): float(s) for s, pa in entries} for _, _, entries in variables]
    best, best_par = -1e300, None
    for order in itertools.permutations(range(4)):
        tot, par = 0.0, {}
        for i, v in enumerate(order):
            U = frozenset(order[:i])
            s, S = max((s, sorted(S)) for S, s in table[v].items() if S <= U)
            tot += s
            par[v] = frozenset(S)
        if tot > best + 1e-9:
            best, best_par = tot, par
    skel = {frozenset((a, b)) for b in best_par for a in best_par[b]}
    vs = {(min(a, c), b, max(a, c)) for b in best_par for a in best_par[b] for c in best_par[b] if a < c and frozenset((a, c)) not in skel}
    return skel, vs


@pytest.mark.parametrize("fig,fn,dag", [("Figure_1", "raw_data_8000.csv", "astar_dag_8000.csv"), ("Figure_2", "raw_data_5000.csv", "astar_dag_5000.csv")])
def test_downstream_mec_from_gpu_pss_matches_reference_golden(pkg, orc, data_dir, tmp_path, fig, fn, dag):
    """BASELINE configs[1]: the .pss written by the GPU `score` binary yields an optimal DAG in the Markov equivalence
    class of the reference's published astar_dag_*.csv ((i,j)=1 means j -> i)."""
    exe = os.path.join(ROOT, "urlearning-cpp_b200", "score")
    out = str(tmp_path / "gpu.pss")
    subprocess.check_call([exe, os.path.join(data_dir, fig, fn), out, "-k", os.path.join(data_dir, "skeleton4_ones.csv"), "-f", "cBIC", "--lambda=2", "--quiet"],
                          stdout=subprocess.DEVNULL)
    _, variables = orc.parse_pss(out)
    skel, vs = _mec_of_best_dag(variables)
    g = np.loadtxt(os.path.join(data_dir, fig, dag), delimiter=",")
    gpar = {i: frozenset(j for j in range(4) if g[i, j] == 1) for i in range(4)}
    gskel = {frozenset((a, b)) for b in gpar for a in gpar[b]}
    gvs = {(min(a, c), b, max(a, c)) for b in gpar for a in gpar[b] for c in gpar[b] if a < c and frozenset((a, c)) not in gskel}
    assert skel == gskel and vs == gvs


def test_score_binary_cbic_matches_oracle(pkg, orc, data_dir, tmp_path):
    exe = os.path.join(ROOT, "urlearning-cpp_b200", "score")
    skel = os.path.join(data_dir, "skeleton4_ones.csv")
    for fig, fn in (("Figure_1", "raw_data_8000.csv"), ("Figure_2", "raw_data_5000.csv")):
        inp = os.path.join(data_dir, fig, fn)
        out, ref = str(tmp_path / "gpu.pss"), str(tmp_path / "ref.pss")
        subprocess.check_call([exe, inp, out, "-k", skel, "-f", "cBIC", "--lambda=2", "--quiet"], stdout=subprocess.DEVNULL)
        orc.score_file(inp, ref, "cBIC", skeleton=skel, lam=2.0)
        mg, vg = orc.parse_pss(out)
        mr, vr = orc.parse_pss(ref)
        assert mg == mr
        for (ng, ag, eg), (nr, ar, er) in zip(vg, vr):
            assert (ng, ag) == (nr, ar)
            assert [e[1] for e in eg] == [e[1] for e in er]
            for a, b in zip(eg, er):
                assert abs(float(a[0]) - float(b[0])) <= 1e-6 * max(1.0, abs(float(b[0])))


def test_row_sharded_gram_protocol(pkg, orc):
    """config-5 protocol on one GPU: two engines hold two row shards; moments and partial Grams are combined in rank
    order exactly as bench.py does over NCCL; the result equals the one-shot Gram to rounding and scores agree."""
    x, _ = pkg.datagen.linear_gaussian_sem(p=40, n=30011, seed=8)
    x[5] += 250.0
    one = pkg.Engine(0)
    one.set_continuous(x)
    g1 = one.gram()
    cuts = [0, 12000, 30011]
    engs = [pkg.Engine(0) for _ in range(2)]
    for e, a, b in zip(engs, cuts[:-1], cuts[1:]):
        e.shard_begin(x[:, a:b].copy())
    n = x.shape[1]
    s1 = sum(e.shard_moments(None)[0] for e in engs)
    mean = s1 / n
    parts = [e.shard_moments(mean) for e in engs]
    S1 = parts[0][0] + parts[1][0]
    S2 = parts[0][1] + parts[1][1]
    dev = np.sqrt((S2 - S1 * S1 / n) / (n - 1.0))
    for e in engs:
        e.shard_finish(mean, dev, n)
    g = engs[0].gram() + engs[1].gram()
    assert np.max(np.abs(g - g1)) <= 1e-12 * n
    ref = orc.gram(orc.standardise(x))
    assert np.max(np.abs(g - ref)) <= 1e-12 * n
    for e in engs:
        e.set_gram(g, n)
    a = engs[0].score_variable(7, (1 << 12) - 1, 5, pkg.CBIC, lam=2.0).fetch()
    b = engs[1].score_variable(7, (1 << 12) - 1, 5, pkg.CBIC, lam=2.0).fetch()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    c = one.score_variable(7, (1 << 12) - 1, 5, pkg.CBIC, lam=2.0).fetch()
    assert np.array_equal(a[0], c[0]) and np.all(ulp_diff(a[1], c[1]) <= 1)
    for e in engs + [one]:
        e.close()


def test_wide_masks_p_above_64(pkg, orc, cbic_engine):
    """p = 90 > 63 (unrepresentable in the reference): two-word masks; checked against the oracle on the
    relabelled sub-problem of the variable's candidates"""
    engine = cbic_engine
    p, n = 90, 4000
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=12, mean_indegree=1.5)
    engine.set_continuous(x)
    rng = np.random.default_rng(4)
    for v in (3, 70, 89):
        cand = sorted(int(i) for i in rng.choice([i for i in range(p) if i != v], size=9, replace=False))
        nb = sum(1 << i for i in cand) | (1 << v)
        res = engine.score_variable(v, nb, 9, pkg.CBIC, lam=2.0)
        masks, scores = res.fetch()
        assert masks.shape[1] == 2
        sub = sorted(cand + [v])
        xs = x[sub]
        vs = sub.index(v)
        z = orc.standardise(xs)
        om = orc.enumerate_sets(vs, (1 << len(sub)) - 1, len(sub), 9)
        full_of = lambda m: sum(1 << sub[i] for i in range(len(sub)) if (int(m) >> i) & 1)
        # every set's float32 the_score from the engine (within 1 ulp of the oracle's residual form) ...
        r0 = engine.score_variable(v, nb, 9, pkg.CBIC, lam=2.0, flags=pkg.CBIC_NO_ACCEPT)
        m0, neg0 = r0.fetch()
        r0.free()
        gpu_ts = {pkg.words_to_mask(row): -s for row, s in zip(m0, neg0)}
        assert len(gpu_ts) == len(om)
        ts = np.array([gpu_ts[full_of(m)] for m in om], dtype=np.float32)
        ref_ts = np.array([np.float32(orc.cbic_residual(z, vs, int(m), 2.0)) for m in om], dtype=np.float32)
        assert np.all(ulp_diff(ts, ref_ts) <= 1)
        # ... and, given those scores, the acceptance decisions and stored values agree EXACTLY with the oracle's recursion
        stored, val = orc.cbic_accept(vs, len(sub), om, ts)
        want = {full_of(m): s for m, s, st in zip(om, val, stored) if st}
        got = {pkg.words_to_mask(row): s for row, s in zip(masks, scores)}
        assert set(got) == set(want)
        assert all(np.float32(got[k]).view(np.uint32) == np.float32(want[k]).view(np.uint32) for k in got)
        s1, ts64 = engine.score_one(v, sum(1 << i for i in cand[:3]), pkg.CBIC, 2.0)
        r = orc.cbic_residual(z, vs, sum(1 << sub.index(i) for i in cand[:3]), 2.0)
        assert abs(ts64 - r) <= TOL * max(1.0, abs(r))


def test_collinear_columns_pivot_guard(pkg, orc, cbic_engine):
    """duplicate and linearly dependent columns: the reference's arma::solve falls back to a least-squares
    solution for the rank-deficient normal equations (BIC_OLS.cpp:313-315), i.e. RSS(S + redundant column) = RSS(S) and
    only the penalty grows.  The engine's Schur sweeps skip a pivot that has lost all its digits instead of dividing by
    it: no Inf/NaN reaches the cache, and the scores obey exactly that identity."""
    engine = cbic_engine
    p, n, lam = 12, 3000, 2.0
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=44)
    x[5] = x[2]                       # exact duplicate
    x[7] = 2.0 * x[1] - 3.0 * x[4]    # exact linear combination
    engine.set_continuous(x)
    z = orc.standardise(x)
    v = 10
    nb = (1 << p) - 1
    res = engine.score_variable(v, nb, 6, pkg.CBIC, lam=lam, flags=pkg.CBIC_NO_ACCEPT)
    masks, neg = res.fetch()
    res.free()
    ts = {int(m): -float(s) for m, s in zip(masks[:, 0], neg)}
    assert all(np.isfinite(t) for t in ts.values())
    pen = lam * np.log(n)
    checked = 0
    for m, t in ts.items():
        k = bin(m).count("1")
        if k >= 6:
            continue
        # a set containing 2 (or 5): adding the twin changes nothing but the penalty
        if (m >> 2) & 1 and not (m >> 5) & 1:
            assert abs(ts[m | (1 << 5)] - (t + pen)) <= 1e-6 * max(1.0, abs(t)), (m, t, ts[m | (1 << 5)])
            checked += 1
        # a set containing 1 and 4: adding 7 = 2 x1 - 3 x4 changes nothing but the penalty
        if (m >> 1) & 1 and (m >> 4) & 1 and not (m >> 7) & 1 and not ((m >> 2) & 1 and (m >> 5) & 1):
            assert abs(ts[m | (1 << 7)] - (t + pen)) <= 1e-6 * max(1.0, abs(t)), (m, t, ts[m | (1 << 7)])
            checked += 1
        # sets free of redundancy still match the oracle's residual form
        if not ((m >> 2) & 1 and (m >> 5) & 1) and not ((m >> 1) & 1 and (m >> 4) & 1 and (m >> 7) & 1) and checked % 7 == 0:
            r = orc.cbic_residual(z, v, m, lam)
            assert abs(t - r) <= 1e-5 * max(1.0, abs(r))
    assert checked > 200
    # filters run on such data without NaNs poisoning the DP
    r2 = engine.score_variable(v, nb, 6, pkg.CBIC, lam=lam, flags=pkg.PRUNE_DOMINATED)
    m2, s2 = r2.fetch()
    r2.free()
    assert len(s2) > 0 and np.all(np.isfinite(s2))


@pytest.mark.parametrize("p,n,K,seed", [(9, 2000, 8, 1), (10, 1500, 9, 2), (12, 2500, 6, 3), (11, 3000, 10, 4)])
def test_literal_zero_acceptance_matches_the_as_written_recursion(pkg, orc, cbic_engine, p, n, K, seed):
    """URLGPU_CBIC_ACCEPT_LITERAL: find_best_subset_score as written (BIC_OLS.cpp:125-172, zero-filled uvec, XOR clear),
    emulated per parent set on the device, against the oracle's emulation of the same lines: identical stored keys and values
    given the engine's float32 scores — for variable 0 (the toggled variable is the child), for variables that have variable 0
    as a candidate, and for a family that does not contain it."""
    engine = cbic_engine
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=100 + seed)
    engine.set_continuous(x)
    differs = 0
    for v, nb in [(0, (1 << p) - 1), (p // 2, (1 << p) - 1), (p - 1, (1 << p) - 1), (3, ((1 << p) - 1) & ~1)]:
        om = orc.enumerate_sets(v, nb, p, K)
        order = orc.canonical_order(om)
        om = om[order]
        r0 = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.CBIC_NO_ACCEPT)
        m0, neg = r0.fetch()
        r0.free()
        assert np.array_equal(m0[:, 0], om)
        ts = -neg
        stored_lit, val_lit = orc.cbic_accept(v, p, om, ts, mode=1)
        stored_cl, _ = orc.cbic_accept(v, p, om, ts, mode=0)
        differs += int((stored_lit != stored_cl).sum())
        r = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.CBIC_ACCEPT_LITERAL)
        m, s = r.fetch()
        r.free()
        assert np.array_equal(m[:, 0], om[stored_lit])
        assert np.array_equal(s.view(np.uint32), val_lit[stored_lit].view(np.uint32))
        # with the prune on top
        keep = orc.prune(om[stored_lit], val_lit[stored_lit], K)
        r = engine.score_variable(v, nb, K, pkg.CBIC, lam=2.0, flags=pkg.CBIC_ACCEPT_LITERAL | pkg.PRUNE_DOMINATED)
        m, s = r.fetch()
        r.free()
        assert np.array_equal(m[:, 0], om[stored_lit][keep])
    assert differs >= 0  # informational: how many keys the two acceptance modes disagree on for this data set
    with pytest.raises(pkg.UrlGpuError, match="at most 12 parents"):
        x2, _ = pkg.datagen.linear_gaussian_sem(p=15, n=500, seed=5)
        engine.set_continuous(x2)
        engine.score_variable(1, (1 << 15) - 1, 14, pkg.CBIC, lam=2.0, flags=pkg.CBIC_ACCEPT_LITERAL)


def test_score_binary_literal_zero_matches_oracle(pkg, orc, tmp_path):
    """`score --accept literal-zero` against the oracle's `.pss` written in its literal-zero mode (same keys, scores to %f noise)"""
    exe = os.path.join(ROOT, "urlearning-cpp_b200", "score")
    p, n = 9, 2000
    x, _ = pkg.datagen.linear_gaussian_sem(p=p, n=n, seed=77)
    inp, out, ref = str(tmp_path / "x.csv"), str(tmp_path / "gpu.pss"), str(tmp_path / "ref.pss")
    pkg.datagen.write_csv(inp, x, fmt="%.17g")
    subprocess.check_call([exe, inp, out, "-f", "cBIC", "--lambda=2", "--accept", "literal-zero", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "cBIC", lam=2.0, accept_mode=1)
    mg, vg = orc.parse_pss(out)
    mr, vr = orc.parse_pss(ref)
    assert mg == mr
    mismatched = 0
    for (ng, ag, eg), (nr, ar, er) in zip(vg, vr):
        assert (ng, ag) == (nr, ar)
        kg, kr = [tuple(e[1]) for e in eg], [tuple(e[1]) for e in er]
        if kg != kr:  # a float32 score on a rounding boundary can flip one acceptance: the engine and the oracle form the_score differently
            mismatched += len(set(kg) ^ set(kr))
    assert mismatched <= 2
