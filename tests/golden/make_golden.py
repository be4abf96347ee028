"""Regenerates the frozen known answers under tests/golden/ from the CPU oracle.

The reference ships no golden .pss; the fixtures below freeze (a) the survey-time probes of the
reference's formulas and (b) the oracle's own output on the reference's data files, so that any later change of the
oracle's arithmetic is caught.  When oracle/_ref (the reference's own sources compiled against shim headers) is
available, tests/test_ref_pin.py additionally checks the oracle against the real reference code.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as orc  # noqa: E402

DATA = os.path.join(os.path.dirname(HERE), "data")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    out = {}
    tmp = "/tmp/_golden.pss"
    # paths inside the .pss header are relative to tests/ so the hash is location independent
    os.chdir(os.path.dirname(HERE))
    n = orc.score_file("data/hepatitis.clean.csv", tmp, "BIC", has_header=True)
    out["hepatitis_bic"] = {"scores": n, "sha256": sha(tmp)}
    n = orc.score_file("data/hepatitis.clean.csv", tmp, "BIC", has_header=True, prune=True)
    out["hepatitis_bic_pruned"] = {"scores": n, "sha256": sha(tmp)}
    for fig, fn in (("Figure_1", "raw_data_8000.csv"), ("Figure_2", "raw_data_5000.csv")):
        n = orc.score_file(f"data/{fig}/{fn}", tmp, "cBIC", skeleton="data/skeleton4_ones.csv", lam=2.0)
        meta, variables = orc.parse_pss(tmp)
        out[f"{fig}_cbic"] = {"scores": n, "lines": [[name, arity, entries] for name, arity, entries in variables]}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
