"""CPU suite (no GPU): the oracle against the reference's fixtures and the frozen known answers, the host-side
logic, the .pss grammar, and the C-ABI library's symbol table."""
import ctypes
import hashlib
import itertools
import json
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


# ------------------------------------------------------------------------------------------- inputs

def test_hepatitis_table(orc, data_dir):
    t = orc.Table(os.path.join(data_dir, "hepatitis.clean.csv"), has_header=True)
    assert (t.n, t.p) == (80, 20)
    assert t.names[0] == "CLASS" and t.names[-1] == "HISTOLOGY"
    assert t.card.tolist() == [2] * 14 + [4, 4, 3, 3, 3, 2]
    codes = t.codes()
    # value index = order of first appearance of the value string (variable.h:43-48): first record is all zeros
    assert (codes[:, 0] == 0).all()
    assert orc.effective_max_parents(0, 20, 80, True) == 3  # (int)ln(160/ln 80)
    assert orc.effective_max_parents(0, 4, 5000, False) == 3
    assert orc.effective_max_parents(12, 60, 1_000_000, True) == 11


def test_csv_reader_quirks(orc, tmp_path):
    f = tmp_path / "q.csv"
    f.write_text("a,b,c\n 1,,2,x \n1.0,2,y\n1,2,x\n")
    t = orc.Table(str(f), has_header=True)
    # token_compress_on merges the empty field; "1" and "1.0" are different values (record.h:35-39)
    assert (t.n, t.p) == (3, 3)
    assert t.card.tolist() == [2, 1, 2]
    g = tmp_path / "ragged.csv"
    g.write_text("1,2,3\n1,2\n")
    with pytest.raises(RuntimeError, match="ragged"):
        orc.Table(str(g))
    with pytest.raises(RuntimeError, match="cannot open"):
        orc.Table(str(tmp_path / "missing.csv"))


def test_skeleton_readers(orc, tmp_path, data_dir):
    edges, init = orc.read_skeleton(os.path.join(data_dir, "skeleton4_ones.csv"), 4)
    assert init and edges == [15] * 4
    m = tmp_path / "m.csv"
    m.write_text("0,TRUE,0,0\n0,0,1,0\n0,0,0,0.9\n0,0,0,0\n")
    edges, init = orc.read_skeleton(str(m), 4)
    # symmetrised; `abs(atof(x)) > 0.05` binds ::abs(int) in the reference as compiled by GCC, so 0.9 is NOT an edge
    # (skeleton.cpp:91; pinned against the reference's own code in test_ref_pin.py)
    assert edges == [0b0010, 0b0101, 0b0010, 0]
    assert orc.two_hop(edges, 4, True, 0) == 0b0111  # N(0) | N(1)
    assert orc.two_hop(edges, 4, True, 3) == 0
    assert orc.two_hop(edges, 4, False, 3) == 0b1111  # uninitialised skeleton: all ones (skeleton.hpp:57-60)
    a = tmp_path / "s.arc"
    a.write_text("V_1,V_3\nV_2,V_3\n")
    edges, init = orc.read_skeleton(str(a), 3)
    assert edges == [0b100, 0b100, 0b011]
    with pytest.raises(RuntimeError):
        orc.read_skeleton(str(tmp_path / "nope.csv"), 4)


def test_enumeration_is_gosper_colex(orc):
    masks = orc.enumerate_sets(2, 0b11111, 5, 3)
    # empty first, then layers; within a layer increasing numeric order of the compact mask; sets with v skipped
    assert masks[0] == 0
    assert all(not (int(m) >> 2) & 1 for m in masks)
    layers = [bin(int(m)).count("1") for m in masks]
    assert layers == sorted(layers)
    for l in (1, 2, 3):
        ms = [int(m) for m in masks if bin(int(m)).count("1") == l]
        assert ms == sorted(ms) and len(ms) == math.comb(4, l)
    assert len(masks) == 1 + 4 + 6 + 4
    # neighbours not containing v, fewer neighbours than the limit
    assert [int(m) for m in orc.enumerate_sets(0, 0b0110, 4, 3)] == [0, 2, 4, 6]


# ------------------------------------------------------------------------------------------- BIC

def test_bic_known_answers(orc, data_dir):
    t = orc.Table(os.path.join(data_dir, "hepatitis.clean.csv"), has_header=True)
    codes = t.codes()
    s, ll = orc.bic_score(codes, t.card, 0, 0)
    # SURVEY.md §3.4 Q4 probe: exact -37.694397, Q4 rule prints -37.694389
    assert "%f" % s == "-37.694389"
    n0 = int((codes[0] == 0).sum())
    exact = n0 * math.log(n0) + (80 - n0) * math.log(80 - n0) - 80 * math.log(80) - math.log(80) / 2
    assert abs(exact - (-37.694397)) < 1e-6 and abs(s - exact) < 2e-5
    assert "%f" % orc.bic_score(codes, t.card, 0, 0b10)[0] == "-39.762127"
    assert "%f" % orc.bic_score(codes, t.card, 0, 0b100)[0] == "-37.771343"


def test_bic_literal_float32_vs_q4_rule(orc, data_dir):
    """the Q4 rule and a literal float32 running sum (the reference's arithmetic) agree to float32 noise"""
    t = orc.Table(os.path.join(data_dir, "hepatitis.clean.csv"), has_header=True)
    codes = t.codes()
    masks = orc.enumerate_sets(3, (1 << 20) - 1, 20, 3)
    a = orc.bic_score_many(codes, t.card, 3, masks, mode=0)
    b = orc.bic_score_many(codes, t.card, 3, masks, mode=1)
    assert np.max(np.abs(a - b) / np.abs(a)) < 2e-6  # SURVEY: up to 1.06e-6 relative on hepatitis
    assert (a < 0).all()


def test_bic_counts_against_numpy(orc):
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 3, size=(6, 500)).astype(np.uint8)
    card = np.full(6, 3, dtype=np.int32)
    cnt = orc.bic_counts(codes, card, 4, 0b100011)
    ref = np.zeros(3 ** 4, dtype=np.int32)
    idx = codes[4].astype(int) + 3 * (codes[0] + 3 * codes[1].astype(int) + 9 * codes[5].astype(int))
    np.add.at(ref, idx, 1)
    assert np.array_equal(cnt, ref) and cnt.sum() == 500


def test_hepatitis_pss_frozen(orc, tmp_path):
    os.chdir(os.path.join(ROOT, "tests"))
    out = str(tmp_path / "h.pss")
    assert orc.score_file("data/hepatitis.clean.csv", out, "BIC", has_header=True) == GOLD["hepatitis_bic"]["scores"] == 23200
    assert sha(out) == GOLD["hepatitis_bic"]["sha256"]
    # threads only change who computes what (score_main.cpp:136-139)
    out2 = str(tmp_path / "h2.pss")
    orc.score_file("data/hepatitis.clean.csv", out2, "BIC", has_header=True, threads=3)
    assert sha(out2) == GOLD["hepatitis_bic"]["sha256"]
    assert orc.score_file("data/hepatitis.clean.csv", out, "BIC", has_header=True, prune=True) == GOLD["hepatitis_bic_pruned"]["scores"]
    assert sha(out) == GOLD["hepatitis_bic_pruned"]["sha256"]


def test_pss_grammar_roundtrip(orc, tmp_path):
    os.chdir(os.path.join(ROOT, "tests"))
    out = str(tmp_path / "h.pss")
    orc.score_file("data/hepatitis.clean.csv", out, "BIC", has_header=True)
    text = open(out).read()
    assert text.startswith("META pss_version = 0.1\nMETA input_file=data/hepatitis.clean.csv\nMETA num_records=80\n"
                           "META parent_limit=3\nMETA score_type=bic\nMETA ess=1\n\nVAR CLASS\nMETA arity=2\n-37.694389 \n-39.762127 AGE \n")
    meta, variables = orc.parse_pss(out)
    assert len(variables) == 20 and all(len(v[2]) == 1160 for v in variables)
    for line in text.split("\n"):
        if line and not line.startswith(("META", "VAR")):
            assert re.fullmatch(r"-?\d+\.\d{6} (\S+ )*", line)


# ------------------------------------------------------------------------------------------- cBIC

def test_cbic_forms_agree(orc, data_dir):
    t = orc.Table(os.path.join(data_dir, "Figure_1", "raw_data_8000.csv"))
    x = t.values()
    z = orc.standardise(x)
    assert np.allclose(z.mean(axis=1), 0, atol=1e-12) and np.allclose(z.var(axis=1, ddof=1), 1, rtol=1e-12)
    g = orc.gram(z)
    assert np.allclose(np.diag(g), 4999.0, rtol=1e-12)  # sample std: G_vv = n-1
    for v in range(4):
        for m in orc.enumerate_sets(v, 15, 4, 3):
            a = orc.cbic_residual(z, v, int(m), 2.0)
            b = orc.cbic_gram(g, 5000, v, int(m), 2.0)
            assert abs(a - b) <= 1e-9 * max(1.0, abs(a))
            if m == 0:
                assert a == 0.0
    # numpy cross-check of one regression
    v, pa = 3, [0, 2]
    beta, *_ = np.linalg.lstsq(z[pa].T, z[v], rcond=None)
    rss = float(((z[v] - beta @ z[pa]) ** 2).sum())
    ref = 5000 * math.log(rss / 5000) + 2.0 * math.log(5000) * 2
    assert abs(orc.cbic_residual(z, v, 0b101, 2.0) - ref) <= 1e-9 * abs(ref)


def _best_dag_mec(p, variables):
    """exhaustive search over DAGs on p<=4 variables; local score = best stored subset (sparse_parent_list.cpp:44-55).
    Returns (skeleton, v-structures) of the optimum = its Markov equivalence class."""
    names = [v[0] for v in variables]
    table = []
    for name, arity, entries in variables:
        d = {}
        for s, pa in entries:
            d[frozenset(names.index(q) for q in pa)] = float(s)
        table.append(d)

    def local(v, U):
        return max(s for S, s in table[v].items() if S <= U)

    best, best_dag = -1e300, None
    for order in itertools.permutations(range(p)):
        tot, dag = 0.0, []
        for i, v in enumerate(order):
            U = frozenset(order[:i])
            s, S = max((s, S) for S, s in table[v].items() if S <= U)
            tot += s
            dag.append((v, S))
        if tot > best + 1e-9:
            best, best_dag = tot, dag
    parents = {v: S for v, S in best_dag}
    skel = {frozenset((a, b)) for b in parents for a in parents[b]}
    vstruct = {(min(a, c), b, max(a, c)) for b in parents for a in parents[b] for c in parents[b] if a < c and frozenset((a, c)) not in skel}
    return skel, vstruct


@pytest.mark.parametrize("fig,fn,dag", [("Figure_1", "raw_data_8000.csv", "astar_dag_8000.csv"),
                                        ("Figure_2", "raw_data_5000.csv", "astar_dag_5000.csv")])
def test_figures_mec_matches_reference_golden(orc, data_dir, tmp_path, fig, fn, dag):
    """end-to-end pin against the reference's own published result: cBIC lambda=2 scores -> optimal DAG is in the
    Markov equivalence class of the reference's astar_dag_*.csv (entry (i,j)=1 means j -> i, Figure_1/README.txt)."""
    os.chdir(os.path.join(ROOT, "tests"))
    out = str(tmp_path / "f.pss")
    n = orc.score_file(f"data/{fig}/{fn}", out, "cBIC", skeleton="data/skeleton4_ones.csv", lam=2.0)
    meta, variables = orc.parse_pss(out)
    assert n == GOLD[f"{fig}_cbic"]["scores"]
    assert [[a, b, [list(e) for e in c]] for a, b, c in variables] == GOLD[f"{fig}_cbic"]["lines"]
    skel, vs = _best_dag_mec(4, variables)
    g = np.loadtxt(os.path.join(data_dir, fig, dag), delimiter=",")
    gpar = {i: frozenset(j for j in range(4) if g[i, j] == 1) for i in range(4)}
    gskel = {frozenset((a, b)) for b in gpar for a in gpar[b]}
    gvs = {(min(a, c), b, max(a, c)) for b in gpar for a in gpar[b] for c in gpar[b] if a < c and frozenset((a, c)) not in gskel}
    assert skel == gskel and vs == gvs


def test_cbic_accept_rules(orc):
    # layers ascending; ts>0 stored negative by the caller, ts==0 dropped, ts<0 stored iff better than every cached subset
    masks = np.array([0, 1, 2, 4, 3, 5, 6, 7], dtype=np.uint64)
    ts = np.array([0, -10, 5, 0, -12, -9, -3, -11], dtype=np.float32)
    stored, val = orc.cbic_accept(3, 4, masks, ts, 0)
    assert stored.tolist() == [True, True, True, False, True, False, True, False]
    assert val[stored].tolist() == [-0.0, 10.0, -5.0, 12.0, 3.0]
    # {0,1,2}: best cached subset is {0,1}=12 >= 11 -> rejected; {1,2}=3 accepted because F({1,2}) = max(0, g{1}=-5, g{2}=F{2}=0) = 0
    s2, v2 = orc.cbic_accept(3, 4, masks, ts, 1)
    assert s2.tolist() == stored.tolist()  # literal-zero mode coincides at p=4


def test_prune_literal(orc):
    masks = np.array([0, 1, 2, 3, 4, 5], dtype=np.uint64)
    scores = np.array([-10, -8, -12, -8, -9, -7.5], dtype=np.float32)
    keep = orc.prune(masks, scores)
    # {1} beats {}, {2} loses to {}, {0,1} ties with its subset {0} and loses, {2'}=4 beats {}, {0,2} beats {0} and {2}
    assert keep.tolist() == [True, True, False, False, True, True]
    assert orc.prune(masks[keep], scores[keep]).all()  # idempotent


# ------------------------------------------------------------------------------------------- product library (no compute)

def test_abi_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "urlgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(urlgpu_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(pkg.ABI_SYMBOLS)
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.UrlGpuError, match="no CPU fallback"):
        pkg.Engine(0)
    assert pkg.device_count() == 0


def test_product_does_not_reference_the_oracle():
    """the product path (package + include/) must not import, link or call anything under oracle/"""
    bad = []
    for base in ("urlearning-cpp_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"oracle_lib|liboracle|orc_[a-z_]+\(|oracle/", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_host_mirrors_match_oracle(pkg, orc):
    rng = np.random.default_rng(5)
    for _ in range(50):
        p = int(rng.integers(2, 30))
        edges = [0] * p
        for _ in range(p):
            a, b = rng.integers(0, p, size=2)
            edges[a] |= 1 << int(b)
            edges[b] |= 1 << int(a)
        v = int(rng.integers(p))
        assert pkg.two_hop_neighbors(edges, p, v) == orc.two_hop(edges, p, True, v)
        assert pkg.two_hop_neighbors(None, p, v) == orc.two_hop(edges, p, False, v)
        n = int(rng.integers(10, 10 ** 7))
        mp = int(rng.integers(0, 40))
        for bic in (True, False):
            assert pkg.effective_max_parents(mp, p, n, bic) == orc.effective_max_parents(mp, p, n, bic)


def test_datagen_is_seeded_and_bounded(pkg):
    a = pkg.datagen.discrete_bn(p=30, n=2000, seed=4)
    b = pkg.datagen.discrete_bn(p=30, n=2000, seed=4)
    assert np.array_equal(a[0], b[0]) and a[2] == b[2]
    codes, card, edges, parents = a
    assert all(len(pa) <= 3 for pa in parents) and codes.max() < 4
    for v in range(30):
        c = bin(pkg.two_hop_neighbors(edges, 30, v) & ~(1 << v)).count("1")
        assert c <= 16
        assert all(codes[v][np.unique(codes[v], return_index=True)[1]].tolist() == sorted(np.unique(codes[v]).tolist()) for _ in [0])
    x, _ = pkg.datagen.linear_gaussian_sem(p=8, n=500, seed=3)
    assert x.shape == (8, 500) and np.isfinite(x).all()
