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
