"""ctypes binding of liburlsearch.so (include/urlsearch.h): the host-side consumers of the `.pss` — reader, sparse parent
list / bitwise best-score structures, static pattern database and A* — restated without Boost
(urlearning-cpp_b200/host/search_host.hpp).  No GPU involved: the search side stays on the host."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liburlsearch.so")
ABI_SYMBOLS = ["urlsearch_open", "urlsearch_close", "urlsearch_last_error", "urlsearch_variable_count", "urlsearch_name", "urlsearch_arity",
               "urlsearch_meta", "urlsearch_entries", "urlsearch_best_scores", "urlsearch_astar", "urlsearch_triplet"]
FLT_MAX = float(np.finfo(np.float32).max)
_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C urlearning-cpp_b200`")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.urlsearch_open.restype = vp
        L.urlsearch_open.argtypes = [C.c_char_p]
        L.urlsearch_close.argtypes = [vp]
        L.urlsearch_close.restype = None
        L.urlsearch_last_error.restype = C.c_char_p
        L.urlsearch_last_error.argtypes = [vp]
        L.urlsearch_variable_count.argtypes = [vp]
        L.urlsearch_name.restype = C.c_char_p
        L.urlsearch_name.argtypes = [vp, i32]
        L.urlsearch_arity.argtypes = [vp, i32]
        L.urlsearch_meta.restype = C.c_char_p
        L.urlsearch_meta.argtypes = [vp, C.c_char_p]
        L.urlsearch_entries.restype = i64
        L.urlsearch_entries.argtypes = [vp, i32, vp, vp, i64]
        L.urlsearch_best_scores.argtypes = [vp, C.c_char_p, i32, vp, i64, vp, vp]
        L.urlsearch_astar.argtypes = [vp, C.c_char_p, i32, C.c_char_p, C.POINTER(C.c_float), vp, C.POINTER(C.c_int)]
        L.urlsearch_triplet.argtypes = [vp, C.c_char_p, i32, C.c_char_p, vp, vp]
        _lib = L
    return _lib


class ScoreCache:
    """scoring::ScoreCache::read of a `.pss` (score_cache.cpp:55-162); scores carry the search side's sign (negated)."""

    def __init__(self, path: str):
        self.lib = load_library()
        self._h = self.lib.urlsearch_open(path.encode())
        if not self._h:
            raise RuntimeError(self.lib.urlsearch_last_error(None).decode())
        self.p = self.lib.urlsearch_variable_count(self._h)
        self.names = [self.lib.urlsearch_name(self._h, v).decode() for v in range(self.p)]
        self.arity = [self.lib.urlsearch_arity(self._h, v) for v in range(self.p)]

    def meta(self, key: str) -> str:
        return self.lib.urlsearch_meta(self._h, key.encode()).decode()

    def entries(self, v: int):
        """(masks uint64 [n], scores float32 [n]) sorted by (score, |S|, mask): the order of the sparse parent list"""
        n = self.lib.urlsearch_entries(self._h, v, None, None, 0)
        masks, scores = np.zeros(n, dtype=np.uint64), np.zeros(n, dtype=np.float32)
        self.lib.urlsearch_entries(self._h, v, masks.ctypes.data, scores.ctypes.data, n)
        return masks, scores

    def best_scores(self, v: int, queries, kind: str = "list"):
        q = np.ascontiguousarray(queries, dtype=np.uint64)
        best, parents = np.zeros(len(q), dtype=np.float32), np.zeros(len(q), dtype=np.uint64)
        if self.lib.urlsearch_best_scores(self._h, kind.encode(), v, q.ctypes.data, len(q), best.ctypes.data, parents.ctypes.data):
            raise RuntimeError(self.lib.urlsearch_last_error(self._h).decode())
        return best, parents

    def astar(self, kind: str = "list", pd_count: int = 2, skeleton: str | None = None):
        """-> (total cost, parents uint64 [p], nodes expanded, components)"""
        cost, nodes = C.c_float(), C.c_int()
        parents = np.zeros(self.p, dtype=np.uint64)
        rc = self.lib.urlsearch_astar(self._h, kind.encode(), pd_count, (skeleton or "").encode(), C.byref(cost), parents.ctypes.data, C.byref(nodes))
        if rc < 0:
            raise RuntimeError(self.lib.urlsearch_last_error(self._h).decode())
        return cost.value, parents, nodes.value, rc

    def triplet(self, skeleton: str, kind: str = "list", pd_count: int = 2):
        """Triplet A* (astar/triplet_astar.cpp) -> (M int32 [p, p] with M[i, j] = 1 iff i -> j, both set for an undirected edge;
        stats dict)"""
        m = np.zeros((self.p, self.p), dtype=np.int32)
        st = np.zeros(5, dtype=np.int32)
        if self.lib.urlsearch_triplet(self._h, kind.encode(), pd_count, skeleton.encode(), m.ctypes.data, st.ctypes.data) < 0:
            raise RuntimeError(self.lib.urlsearch_last_error(self._h).decode())
        return m, dict(zip(("triples", "colliders", "unfaithful_edges", "oriented_by_rules", "nodes_expanded"), (int(x) for x in st)))

    def close(self):
        if self._h:
            self.lib.urlsearch_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
