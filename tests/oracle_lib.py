"""ctypes binding of the CPU oracle (oracle/build/liboracle.so) — test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "build", "liboracle.so")
CLI = os.path.join(ORACLE_DIR, "build", "oracle_score")


class Options(C.Structure):
    _fields_ = [("input", C.c_char_p), ("output", C.c_char_p), ("skeleton", C.c_char_p), ("function", C.c_char_p),
                ("delimiter", C.c_char), ("has_header", C.c_int), ("max_parents", C.c_int), ("lambda_", C.c_double),
                ("threads", C.c_int), ("prune", C.c_int), ("accept_mode", C.c_int), ("bic_mode", C.c_int),
                ("cbic_from_gram", C.c_int), ("ess", C.c_float)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        L.orc_read_csv.restype = vp
        L.orc_read_csv.argtypes = [C.c_char_p, C.c_char, i32]
        L.orc_table_free.argtypes = [vp]
        L.orc_table_n.restype = i64
        L.orc_table_n.argtypes = [vp]
        L.orc_table_p.argtypes = [vp]
        L.orc_table_name.restype = C.c_char_p
        L.orc_table_name.argtypes = [vp, i32]
        L.orc_table_card.argtypes = [vp, vp]
        L.orc_table_codes.argtypes = [vp, vp]
        L.orc_table_values.argtypes = [vp, vp]
        L.orc_last_error.restype = C.c_char_p
        L.orc_read_skeleton.argtypes = [C.c_char_p, i32, vp]
        L.orc_two_hop.restype = u64
        L.orc_two_hop.argtypes = [vp, i32, i32, i32]
        L.orc_effective_max_parents.argtypes = [i32, i32, i64, i32]
        L.orc_enumerate.restype = i64
        L.orc_enumerate.argtypes = [i32, u64, i32, i32, vp, i64]
        L.orc_bic_cells.restype = i64
        L.orc_bic_cells.argtypes = [vp, i32, i32, u64]
        L.orc_bic_counts.argtypes = [vp, i64, i32, vp, i32, u64, vp]
        L.orc_bic_score.argtypes = [vp, i64, i32, vp, i32, u64, i32, C.POINTER(C.c_float), C.POINTER(C.c_double)]
        L.orc_bic_score_many.argtypes = [vp, i64, i32, vp, i32, vp, i64, i32, i32, vp]
        L.orc_log_regret.argtypes = [i64, i32, vp]
        L.orc_fnml_score_many.argtypes = [vp, i64, i32, vp, i32, vp, i64, i32, i32, vp]
        L.orc_bdeu_score_many.argtypes = [vp, i64, i32, vp, i32, C.c_float, vp, i64, i32, i32, vp]
        L.orc_standardise.argtypes = [vp, i64, i32, vp]
        L.orc_gram.argtypes = [vp, i64, i32, vp]
        L.orc_cbic_the_score_residual.restype = C.c_double
        L.orc_cbic_the_score_residual.argtypes = [vp, i64, i32, i32, u64, C.c_double]
        L.orc_cbic_the_score_gram.restype = C.c_double
        L.orc_cbic_the_score_gram.argtypes = [vp, i64, i32, i32, u64, C.c_double]
        L.orc_cbic_accept.argtypes = [i32, i32, vp, vp, i64, i32, vp, vp]
        L.orc_prune.argtypes = [vp, vp, i64, i32, vp]
        L.orc_score_file.restype = i64
        L.orc_score_file.argtypes = [C.POINTER(Options)]
        _lib = L
    return _lib


def err():
    return lib().orc_last_error().decode()


class Table:
    def __init__(self, path, delimiter=",", has_header=False):
        L = lib()
        self.h = L.orc_read_csv(path.encode(), delimiter.encode(), int(has_header))
        if not self.h:
            raise RuntimeError(err())
        self.n = L.orc_table_n(self.h)
        self.p = L.orc_table_p(self.h)
        self.names = [L.orc_table_name(self.h, i).decode() for i in range(self.p)]
        self.card = np.zeros(self.p, dtype=np.int32)
        L.orc_table_card(self.h, self.card.ctypes.data)

    def codes(self):
        out = np.zeros((self.p, self.n), dtype=np.uint8)
        if lib().orc_table_codes(self.h, out.ctypes.data):
            raise RuntimeError(err())
        return out

    def values(self):
        out = np.zeros((self.p, self.n), dtype=np.float64)
        lib().orc_table_values(self.h, out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_table_free(self.h)
            self.h = None


def read_skeleton(path, p):
    edges = np.zeros(p, dtype=np.uint64)
    rc = lib().orc_read_skeleton(path.encode() if path else None, p, edges.ctypes.data)
    if rc < 0:
        raise RuntimeError(err())
    return [int(e) for e in edges], bool(rc)


def two_hop(edges, p, initialised, v):
    e = np.asarray(edges, dtype=np.uint64)
    return int(lib().orc_two_hop(e.ctypes.data, p, int(initialised), v))


def effective_max_parents(flag, p, n, is_bic):
    return lib().orc_effective_max_parents(flag, p, n, int(is_bic))


def enumerate_sets(v, neighbors, p, max_parents):
    m = lib().orc_enumerate(v, neighbors, p, max_parents, None, 0)
    if m < 0:
        raise RuntimeError(err())
    out = np.zeros(m, dtype=np.uint64)
    lib().orc_enumerate(v, neighbors, p, max_parents, out.ctypes.data, m)
    return out


def bic_cells(card, v, parents):
    card = np.ascontiguousarray(card, dtype=np.int32)
    return lib().orc_bic_cells(card.ctypes.data, len(card), v, parents)


def bic_counts(codes, card, v, parents):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    card = np.ascontiguousarray(card, dtype=np.int32)
    p, n = codes.shape
    out = np.zeros(bic_cells(card, v, parents), dtype=np.int32)
    if lib().orc_bic_counts(codes.ctypes.data, n, p, card.ctypes.data, v, parents, out.ctypes.data):
        raise RuntimeError(err())
    return out


def bic_score(codes, card, v, parents, mode=0):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    card = np.ascontiguousarray(card, dtype=np.int32)
    p, n = codes.shape
    s, ll = C.c_float(), C.c_double()
    if lib().orc_bic_score(codes.ctypes.data, n, p, card.ctypes.data, v, parents, mode, C.byref(s), C.byref(ll)):
        raise RuntimeError(err())
    return np.float32(s.value), ll.value


def bic_score_many(codes, card, v, masks, mode=0, threads=8):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    card = np.ascontiguousarray(card, dtype=np.int32)
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    p, n = codes.shape
    out = np.zeros(len(masks), dtype=np.float32)
    if lib().orc_bic_score_many(codes.ctypes.data, n, p, card.ctypes.data, v, masks.ctypes.data, len(masks), mode, threads,
                                out.ctypes.data):
        raise RuntimeError(err())
    return out


def log_regret(n_max, r):
    """one row of the reference's regret cache: (float)log(reg(N, r)), N = 0..n_max"""
    out = np.zeros(n_max + 1, dtype=np.float32)
    lib().orc_log_regret(n_max, r, out.ctypes.data)
    return out


def fnml_score_many(codes, card, v, masks, mode=0, threads=8):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    card = np.ascontiguousarray(card, dtype=np.int32)
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    p, n = codes.shape
    out = np.zeros(len(masks), dtype=np.float32)
    if lib().orc_fnml_score_many(codes.ctypes.data, n, p, card.ctypes.data, v, masks.ctypes.data, len(masks), mode, threads,
                                 out.ctypes.data):
        raise RuntimeError(err())
    return out


def bdeu_score_many(codes, card, v, masks, ess=1.0, mode=0, threads=8):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    card = np.ascontiguousarray(card, dtype=np.int32)
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    p, n = codes.shape
    out = np.zeros(len(masks), dtype=np.float32)
    if lib().orc_bdeu_score_many(codes.ctypes.data, n, p, card.ctypes.data, v, ess, masks.ctypes.data, len(masks), mode, threads,
                                 out.ctypes.data):
        raise RuntimeError(err())
    return out


def standardise(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    p, n = x.shape
    z = np.zeros_like(x)
    lib().orc_standardise(x.ctypes.data, n, p, z.ctypes.data)
    return z


def gram(z):
    z = np.ascontiguousarray(z, dtype=np.float64)
    p, n = z.shape
    g = np.zeros((p, p), dtype=np.float64)
    lib().orc_gram(z.ctypes.data, n, p, g.ctypes.data)
    return g


def cbic_residual(z, v, parents, lam):
    z = np.ascontiguousarray(z, dtype=np.float64)
    p, n = z.shape
    return lib().orc_cbic_the_score_residual(z.ctypes.data, n, p, v, parents, lam)


def cbic_gram(g, n, v, parents, lam):
    g = np.ascontiguousarray(g, dtype=np.float64)
    return lib().orc_cbic_the_score_gram(g.ctypes.data, n, g.shape[0], v, parents, lam)


def cbic_accept(v, p, masks, the_scores, mode=0):
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    ts = np.ascontiguousarray(the_scores, dtype=np.float32)
    stored = np.zeros(len(masks), dtype=np.uint8)
    val = np.zeros(len(masks), dtype=np.float32)
    lib().orc_cbic_accept(v, p, masks.ctypes.data, ts.ctypes.data, len(masks), mode, stored.ctypes.data, val.ctypes.data)
    return stored.astype(bool), val


def prune(masks, scores, highest_completed_layer=64):
    masks = np.ascontiguousarray(masks, dtype=np.uint64)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    keep = np.zeros(len(masks), dtype=np.uint8)
    lib().orc_prune(masks.ctypes.data, scores.ctypes.data, len(masks), highest_completed_layer, keep.ctypes.data)
    return keep.astype(bool)


def score_file(input, output, function="BIC", skeleton=None, has_header=False, max_parents=0, lam=0.5, threads=1,
               prune=False, accept_mode=0, bic_mode=0, from_gram=False, ess=0.0):
    o = Options(input.encode(), output.encode(), skeleton.encode() if skeleton else None, function.encode(), b",",
                int(has_header), max_parents, lam, threads, int(prune), accept_mode, bic_mode, int(from_gram), ess)
    n = lib().orc_score_file(C.byref(o))
    if n < 0:
        raise RuntimeError(err())
    return n


def canonical_order(masks):
    """indices sorting masks by (|S|, mask)"""
    masks = [int(m) for m in masks]
    return sorted(range(len(masks)), key=lambda i: (bin(masks[i]).count("1"), masks[i]))


def parse_pss(path):
    """Restatement of ScoreCache::read (score_cache/score_cache.cpp:55-162) for round-trip tests.
    -> (meta dict, [(name, arity, [(score, [parent names])...])...])."""
    meta, variables = {}, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if not line or line.startswith("#"):
                continue
            low = line.lower()
            if low.startswith("var "):
                variables.append([line.split(" ")[1], None, []])
            elif low.startswith("meta"):
                body = line[4:].strip()
                k, v = body.split("=", 1)
                if variables:
                    if "arity" in k:
                        variables[-1][1] = int(v)
                else:
                    meta[k.strip()] = v.strip()
            else:
                tok = [t for t in line.split(" ") if t != ""]
                variables[-1][2].append((tok[0], tok[1:]))
    return meta, variables
