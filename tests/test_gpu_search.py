"""Downstream identity (BASELINE.json north_star: "the downstream astar DAG ... identical"): the `.pss` written by the GPU
`score` binary, read back by the reference's own reader (oracle/_ref/libref_search.so) and searched by A*, against the same
pipeline on the CPU oracle's `.pss`."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")
EXE = os.path.join(ROOT, "urlearning-cpp_b200", "score")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_search.so")


@pytest.fixture(scope="module")
def S():
    return importlib.import_module("urlearning-cpp_b200.search")


def test_hepatitis_dag_from_gpu_pss_identical(S, orc, tmp_path):
    """configs[0]: hepatitis, discrete BIC, score -> .pss -> astar DAG (p = 20): GPU and oracle files give the same entries,
    the same optimal cost and the same DAG, with and without --prune (pruning never changes the optimum)"""
    inp = os.path.join(DATA, "hepatitis.clean.csv")
    gpu, ref, gpu_pruned = str(tmp_path / "gpu.pss"), str(tmp_path / "ref.pss"), str(tmp_path / "gpu_pruned.pss")
    subprocess.check_call([EXE, inp, gpu, "-s", "-f", "BIC", "--quiet"], stdout=subprocess.DEVNULL)
    subprocess.check_call([EXE, inp, gpu_pruned, "-s", "-f", "BIC", "--prune", "--quiet"], stdout=subprocess.DEVNULL)
    orc.score_file(inp, ref, "BIC", has_header=True)
    a, b, c = S.ScoreCache(gpu), S.ScoreCache(ref), S.ScoreCache(gpu_pruned)
    for v in range(20):
        (ma, sa), (mb, sb) = a.entries(v), b.entries(v)
        assert np.array_equal(ma, mb) and np.array_equal(sa.view(np.uint32), sb.view(np.uint32))
    ca, pa, na, _ = a.astar("list")
    cb, pb, nb, _ = b.astar("list")
    cc, pc, _, _ = c.astar("bitwise")
    assert np.float32(ca) == np.float32(cb) and np.array_equal(pa, pb) and na == nb
    assert abs(cc - ca) <= 1e-4 * abs(ca)
    assert np.array_equal(pc, pa)
    # the `astar` binary writes that DAG
    net = str(tmp_path / "net")
    subprocess.check_call([os.path.join(ROOT, "urlearning-cpp_b200", "astar"), gpu, "-n", net, "--quiet"], stdout=subprocess.DEVNULL)
    m = np.loadtxt(net + ".csv", delimiter=",")
    assert m.shape == (20, 20)
    assert all(int(sum(int(m[v, i]) << i for i in range(20))) == int(pa[v]) for v in range(20))


@pytest.mark.parametrize("fig,fn,dag", [("Figure_1", "raw_data_8000.csv", "astar_dag_8000.csv"), ("Figure_2", "raw_data_5000.csv", "astar_dag_5000.csv")])
def test_figures_published_dag_from_gpu_pss(S, tmp_path, fig, fn, dag):
    """configs[1]: cBIC lambda=2 on Figure_1/2: A* on the GPU-written `.pss` returns the reference's published DAG edge for edge"""
    skel = os.path.join(DATA, "skeleton4_ones.csv")
    gpu = str(tmp_path / "gpu.pss")
    subprocess.check_call([EXE, os.path.join(DATA, fig, fn), gpu, "-k", skel, "-f", "cBIC", "--lambda=2", "--quiet"], stdout=subprocess.DEVNULL)
    cost, parents, _, _ = S.ScoreCache(gpu).astar("list", 2, skel)
    want = np.loadtxt(os.path.join(DATA, fig, dag), delimiter=",")
    got = np.array([[(int(parents[v]) >> i) & 1 for i in range(4)] for v in range(4)], dtype=float)
    assert np.array_equal(got, want)


def test_reference_reader_round_trips_the_gpu_pss(pkg, tmp_path):
    """the reference's own ScoreCache::read (compiled from score_cache/score_cache.cpp) parses the GPU-written file into
    exactly the cache the engine holds: every (variable, parent set) once, score = -(float)atof(%f text)"""
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_search.so not built")
    L = C.CDLL(REF_SO)
    L.refs_open.restype = C.c_void_p
    L.refs_open.argtypes = [C.c_char_p]
    L.refs_variable_count.argtypes = [C.c_void_p]
    L.refs_entries.restype = C.c_int64
    L.refs_entries.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    p, n, K = 30, 5000, 4
    codes, card, edges, _ = pkg.datagen.discrete_bn(p=p, n=n, seed=91, window=4, max_indegree=3)
    inp, skel, out = str(tmp_path / "d.csv"), str(tmp_path / "skel.csv"), str(tmp_path / "gpu.pss")
    pkg.datagen.write_csv(inp, codes)
    pkg.datagen.write_skeleton_matrix(skel, edges, p)
    subprocess.check_call([EXE, inp, out, "-k", skel, "-f", "BIC", "-p", str(K), "--prune", "--quiet"], stdout=subprocess.DEVNULL)
    h = L.refs_open(out.encode())
    assert h and L.refs_variable_count(h) == p
    eng = pkg.Engine(0)
    eng.set_discrete(codes, card)
    for v in range(p):
        res = eng.score_variable(v, pkg.two_hop_neighbors(edges, p, v), K, pkg.BIC, flags=pkg.PRUNE_DOMINATED)
        masks, scores = res.fetch()
        res.free()
        cnt = L.refs_entries(h, v, None, None, 0)
        rm, rs = np.zeros(cnt, dtype=np.uint64), np.zeros(cnt, dtype=np.float32)
        L.refs_entries(h, v, rm.ctypes.data, rs.ctypes.data, cnt)
        want = {int(m): np.float32(-float("%f" % float(s))) for m, s in zip(masks[:, 0], scores)}
        got = dict(zip(map(int, rm), rs))
        assert got == want
    eng.close()


def test_device_sparse_parent_graph_equals_host_and_reference(pkg, S, orc, engine, tmp_path):
    """urlgpu_spg_*: SparseParentBitwise::getScore as a batched device look-up: same best score as the host
    restatement (and through it the reference's own list / bitwise / tree structures, tests/test_search.py) for random
    allowed sets, the empty set, the full set and sets with no cached subset; multi-word variable sets as well"""
    inp = os.path.join(DATA, "hepatitis.clean.csv")
    pss = str(tmp_path / "hep.pss")
    subprocess.check_call([EXE, inp, pss, "-s", "-f", "BIC", "--quiet"], stdout=subprocess.DEVNULL)
    cache = S.ScoreCache(pss)
    rng = np.random.default_rng(11)
    for v in (0, 7, 19):
        masks, scores = cache.entries(v)
        perm = rng.permutation(len(masks))       # the build sorts
        spg = engine.sparse_parent_graph(masks[perm], scores[perm], cache.p)
        q = rng.integers(0, 1 << 20, size=3000, dtype=np.uint64) & ~np.uint64(1 << v)
        q[:2] = [0, (1 << 20) - 1 - (1 << v)]
        best, parents, index = spg.query(q)
        hb, hp = cache.best_scores(v, q, "bitwise")
        assert np.array_equal(best.view(np.uint32), hb.view(np.uint32))
        assert np.array_equal(parents[:, 0], hp)            # same deterministic order among equal scores
        assert np.array_equal(masks[index], parents[:, 0])
        spg.free()
    # entries that all need variable 3: queries without it find nothing (FLT_MAX, sparse_parent_bitwise.cpp:104-106)
    m = np.array([0b1000, 0b1010, 0b1001], dtype=np.uint64)
    spg = engine.sparse_parent_graph(m, np.array([5.0, 3.0, 4.0], dtype=np.float32), 6)
    best, parents, index = spg.query(np.array([0b0111, 0b1000, 0b1011], dtype=np.uint64))
    assert best[0] == np.finfo(np.float32).max and index[0] == -1 and parents[0, 0] == 0
    assert list(best[1:]) == [5.0, 3.0] and list(parents[1:, 0]) == [0b1000, 0b1010]
    spg.free()
    # 150 variables (three words), 50000 entries: against a brute-force scan
    p, n = 150, 50000
    ent = np.zeros((n, 3), dtype=np.uint64)
    for i in range(n):
        for b in rng.choice(p, size=int(rng.integers(0, 4)), replace=False):
            ent[i, b >> 6] |= np.uint64(1 << int(b & 63))
    sc = rng.normal(100, 30, size=n).astype(np.float32)
    spg = engine.sparse_parent_graph(ent, sc, p)
    qs = np.zeros((200, 3), dtype=np.uint64)
    for i in range(200):
        for b in rng.choice(p, size=int(rng.integers(0, 120)), replace=False):
            qs[i, b >> 6] |= np.uint64(1 << int(b & 63))
    best, parents, index = spg.query(qs)
    for i in range(200):
        ok = np.all((ent & ~qs[i]) == 0, axis=1)
        want = sc[ok].min() if ok.any() else np.finfo(np.float32).max
        assert best[i] == want
        if ok.any():
            assert np.all((parents[i] & ~qs[i]) == 0)
    spg.free()
