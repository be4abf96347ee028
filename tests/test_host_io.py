"""Host-side input/output of the `score` binary (urlearning-cpp_b200/host/fast_io.hpp), CPU only:
the parallel CSV reader against the oracle's restatement of RecordFile / BayesianNetwork (first-appearance value coding,
base/variable.h:43-48; token compression, base/record.h:35-39) and the exact `%f` formatter against libc."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "urlearning-cpp_b200", "liburlhost.so")


@pytest.fixture(scope="module")
def H():
    L = C.CDLL(LIB)
    L.urlhost_format_score.argtypes = [C.c_float, C.c_char_p]
    L.urlhost_csv_open.restype = C.c_void_p
    L.urlhost_csv_open.argtypes = [C.c_char_p, C.c_char, C.c_int, C.c_int]
    L.urlhost_last_error.restype = C.c_char_p
    L.urlhost_csv_free.argtypes = [C.c_void_p]
    L.urlhost_csv_p.argtypes = [C.c_void_p]
    L.urlhost_csv_n.restype = C.c_int64
    L.urlhost_csv_n.argtypes = [C.c_void_p]
    L.urlhost_csv_cardinality.argtypes = [C.c_void_p, C.c_int]
    L.urlhost_csv_value.restype = C.c_char_p
    L.urlhost_csv_value.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.urlhost_csv_header.restype = C.c_char_p
    L.urlhost_csv_header.argtypes = [C.c_void_p, C.c_int]
    L.urlhost_csv_codes.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return L


def test_format_score_is_libc_percent_f(H):
    libc = C.CDLL(None)
    libc.snprintf.restype = C.c_int
    rng = np.random.default_rng(0)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 0.25, 1e-7, 5e-7, 4.999999e-7, 1.5e-6, 2.5e-6, 123456.78, -962.494873, 8388608.0, 16777216.0, 3.4e38, 1e-45, 1e13, -1e13, 1.7e13,
            float("inf"), float("-inf"), float("nan"), 0.0000005, 0.0000015, 0.0000025, 1234567.0000005]
    vals += list((rng.standard_normal(20000) * 10 ** rng.uniform(-8, 8, size=20000)).astype(np.float32))
    vals += list(np.ldexp(rng.integers(1, 1 << 24, size=20000).astype(np.float64), rng.integers(-30, 2, size=20000)).astype(np.float32))  # exact halves of the 6th digit included
    a, b = C.create_string_buffer(128), C.create_string_buffer(128)
    for v in vals:
        v = float(np.float32(v))
        n1 = H.urlhost_format_score(C.c_float(v), a)
        n2 = libc.snprintf(b, C.c_size_t(128), b"%f", C.c_double(v))
        assert a.value == b.value and n1 == n2, (v, a.value, b.value)


def _check_against_oracle(H, orc, path, has_header, threads):
    t = orc.Table(path, has_header=has_header)
    h = H.urlhost_csv_open(path.encode(), b",", int(has_header), threads)
    assert h, H.urlhost_last_error()
    assert H.urlhost_csv_p(h) == t.p and H.urlhost_csv_n(h) == t.n
    codes = t.codes()
    for j in range(t.p):
        assert H.urlhost_csv_cardinality(h, j) == t.card[j]
        got = np.zeros(t.n, dtype=np.int32)
        H.urlhost_csv_codes(h, j, got.ctypes.data)
        assert np.array_equal(got, codes[j].astype(np.int32))
        if has_header:
            assert H.urlhost_csv_header(h, j).decode() == t.names[j]
    H.urlhost_csv_free(h)


def test_parse_csv_equals_oracle_reader(H, orc, pkg, tmp_path, data_dir):
    _check_against_oracle(H, orc, os.path.join(data_dir, "hepatitis.clean.csv"), True, 1)
    _check_against_oracle(H, orc, os.path.join(data_dir, "hepatitis.clean.csv"), True, 3)
    # big enough to be cut into several chunks; a value that first appears late in some column must still get the next index
    codes, card, _, _ = pkg.datagen.discrete_bn(p=12, n=400000, seed=3)
    codes[5, :300000] = np.minimum(codes[5, :300000], 1)
    path = str(tmp_path / "big.csv")
    pkg.datagen.write_csv(path, codes)
    for threads in (1, 4, 7):
        _check_against_oracle(H, orc, path, False, threads)


def test_parse_csv_token_rules(H, orc, tmp_path):
    """leading/trailing blanks are trimmed, runs of the delimiter count once, value identity is the string ("1" != "1.0")"""
    path = str(tmp_path / "quirks.csv")
    open(path, "w").write("a,b,c\n 1,x,,y \n1.0,x,y\n1,,z,y\n\t2,x,y\r\n")
    _check_against_oracle(H, orc, path, True, 1)
    h = H.urlhost_csv_open(path.encode(), b",", 1, 1)
    assert [H.urlhost_csv_value(h, 0, k).decode() for k in range(H.urlhost_csv_cardinality(h, 0))] == ["1", "1.0", "2"]
    H.urlhost_csv_free(h)
    bad = str(tmp_path / "short.csv")
    open(bad, "w").write("1,2,3\n1,2\n")
    assert not H.urlhost_csv_open(bad.encode(), b",", 0, 1)
    assert b"fewer fields" in H.urlhost_last_error()
