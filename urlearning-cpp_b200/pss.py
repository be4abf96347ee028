"""`.pss` writer of the host side (score_main.cpp:173-203, 383-402), used by the Python driver and the benches.

Grammar: header META lines, then per variable ``VAR <name>``, ``META arity=<r>``, one line
``"%f " + "<parent> "*`` per cache entry, and a blank line.  Entries arrive in canonical (|S|, mask) order."""
from __future__ import annotations

import ctypes

import numpy as np

_libc = ctypes.CDLL(None)
_libc.snprintf.restype = ctypes.c_int


def format_score(x) -> str:
    """C's ``%f`` of the float promoted to double (score_main.cpp:191) — exactly glibc's rounding."""
    buf = ctypes.create_string_buffer(64)
    _libc.snprintf(buf, ctypes.c_size_t(64), b"%f", ctypes.c_double(float(np.float32(x))))
    return buf.value.decode()


def lexical_float(x: float) -> str:
    """boost::lexical_cast<std::string>(float): 9 significant digits (score_main.cpp:388)."""
    return "%.9g" % float(np.float32(x))


def write_pss(path: str, input_file: str, num_records: int, parent_limit: int, score_type: str, names, arities, caches,
              ess: float = 1.0):
    """caches: {variable: (masks uint64 [n, words], scores float32 [n])}."""
    p = len(names)
    with open(path, "w") as f:
        f.write(f"META pss_version = 0.1\nMETA input_file={input_file}\nMETA num_records={num_records}\n")
        f.write(f"META parent_limit={parent_limit}\nMETA score_type={score_type.lower()}\nMETA ess={lexical_float(ess)}\n\n")
        for v in range(p):
            masks, scores = caches[v]
            f.write(f"VAR {names[v]}\nMETA arity={int(arities[v])}\n")
            lines = []
            for row, s in zip(masks, scores):
                m = 0
                for w, x in enumerate(np.atleast_1d(row)):
                    m |= int(x) << (64 * w)
                parents = "".join(names[q] + " " for q in range(p) if (m >> q) & 1)
                lines.append(format_score(s) + " " + parents + "\n")
            f.write("".join(lines))
            f.write("\n")
