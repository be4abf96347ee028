"""ctypes binding of oracle/_ref/libref_bic.so — the REFERENCE's own discrete-BIC sources compiled against shim
headers (oracle/ref.mk).  Test infrastructure only; absent when the reference was not available at build time."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_bic.so")


def available():
    return os.path.exists(LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        L.ref_open.restype = vp
        L.ref_open.argtypes = [C.c_char_p, C.c_char, i32, i32]
        for f in ("ref_p", "ref_n"):
            getattr(L, f).argtypes = [vp]
        L.ref_cardinality.argtypes = [vp, i32]
        L.ref_name.restype = C.c_char_p
        L.ref_name.argtypes = [vp, i32]
        L.ref_code.argtypes = [vp, i32, i32]
        L.ref_calculate_score.restype = C.c_float
        L.ref_calculate_score.argtypes = [vp, i32, u64]
        L.ref_fnml_score.restype = C.c_float
        L.ref_fnml_score.argtypes = [vp, i32, u64]
        L.ref_regret.restype = C.c_float
        L.ref_regret.argtypes = [vp, i32, i32]
        L.ref_bdeu_score.restype = C.c_float
        L.ref_bdeu_score.argtypes = [vp, i32, u64, C.c_float]
        L.ref_contab.restype = i64
        L.ref_contab.argtypes = [vp, u64, vp, i64]
        L.ref_score_variable.restype = i64
        L.ref_score_variable.argtypes = [vp, i32, u64, i32, i32, vp, vp, i64]
        L.ref_score_all.restype = i64
        L.ref_score_all.argtypes = [vp, vp, i32, i32]
        L.ref_read_skeleton.argtypes = [C.c_char_p, i32, vp]
        _lib = L
    return _lib


class Reference:
    """score_main.cpp:283-347 for -f BIC: RecordFile -> BayesianNetwork -> ADTree -> LLC -> BICScoringFunction."""

    def __init__(self, csv, has_header=False, delimiter=",", r_min=5):
        # the reference prints progress to stdout (printf); silence is not needed for correctness
        self.h = lib().ref_open(csv.encode(), delimiter.encode(), int(has_header), r_min)
        if not self.h:
            raise RuntimeError("reference could not read " + csv)
        self.p, self.n = lib().ref_p(self.h), lib().ref_n(self.h)
        self.card = np.array([lib().ref_cardinality(self.h, v) for v in range(self.p)], dtype=np.int32)
        self.names = [lib().ref_name(self.h, v).decode() for v in range(self.p)]

    def codes(self):
        return np.array([[lib().ref_code(self.h, v, r) for r in range(self.n)] for v in range(self.p)], dtype=np.uint8)

    def calculate_score(self, v, parents):
        return np.float32(lib().ref_calculate_score(self.h, v, parents))

    def fnml_score(self, v, parents):
        return np.float32(lib().ref_fnml_score(self.h, v, parents))

    def regret(self, arity, N):
        return np.float32(lib().ref_regret(self.h, arity, N))

    def bdeu_score(self, v, parents, ess=1.0):
        return np.float32(lib().ref_bdeu_score(self.h, v, parents, ess))

    def contab(self, variables):
        cells = lib().ref_contab(self.h, variables, None, 0)
        out = np.zeros(cells, dtype=np.int32)
        lib().ref_contab(self.h, variables, out.ctypes.data, cells)
        return out

    def score_variable(self, v, neighbors, max_parents, prune=False):
        m = lib().ref_score_variable(self.h, v, neighbors, max_parents, int(prune), None, None, 0)
        masks = np.zeros(m, dtype=np.uint64)
        scores = np.zeros(m, dtype=np.float32)
        lib().ref_score_variable(self.h, v, neighbors, max_parents, int(prune), masks.ctypes.data, scores.ctypes.data, m)
        return masks, scores

    def score_all(self, neighbors, max_parents, threads):
        nb = np.ascontiguousarray(neighbors, dtype=np.uint64)
        return lib().ref_score_all(self.h, nb.ctypes.data, max_parents, threads)


def read_skeleton(path, p):
    edges = np.zeros(p, dtype=np.uint64)
    if lib().ref_read_skeleton(path.encode(), p, edges.ctypes.data):
        raise RuntimeError("reference skeleton reader failed")
    return [int(e) for e in edges]
