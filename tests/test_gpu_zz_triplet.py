"""The whole published pipeline of Figure_1 / Figure_2 with the GPU in it: raw data -> GPU `score` (cBIC lambda=2) -> `.pss`
-> Triplet A* (host/triplet_host.hpp) -> the reference's published Markov equivalence class
(triplet_data/Figure_*/triplet_mec_*.csv), edge for edge.  (The file sorts last on purpose: everything else runs first.)"""
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")
SCORE = os.path.join(ROOT, "urlearning-cpp_b200", "score")


@pytest.mark.parametrize("fig,n,skeleton", [("Figure_1", 8000, "skeleton4_ones.csv"), ("Figure_2", 5000, "skeleton4_cycle.csv")])
def test_published_mec_from_gpu_pss(tmp_path, fig, n, skeleton):
    S = importlib.import_module("urlearning-cpp_b200.search")
    skel = os.path.join(DATA, skeleton)
    gpu = str(tmp_path / "gpu.pss")
    subprocess.check_call([SCORE, os.path.join(DATA, fig, f"raw_data_{n}.csv"), gpu, "-k", skel, "-f", "cBIC", "--lambda=2", "--quiet"], stdout=subprocess.DEVNULL)
    want = np.loadtxt(os.path.join(DATA, fig, f"triplet_mec_{n}.csv"), delimiter=",", dtype=np.int32)
    cache = S.ScoreCache(gpu)
    got, stats = cache.triplet(skel)
    cache.close()
    assert np.array_equal(got, want), (got, stats)
