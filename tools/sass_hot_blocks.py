#!/usr/bin/env python
"""Summarises `ncu -i report --page source --csv` (SASS view, one section per captured launch): for the chosen launches, the
straight-line SASS blocks that carry most of the executed warp instructions, with their lane occupancy and stall samples.
Usage: python tools/sass_hot_blocks.py source.csv [kernel-substring ...] > profiles/rNN_k1_sass_hot_blocks.txt"""
import csv
import sys


def sections(path):
    cur = None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if row and row[0] == "Kernel Name":
                if cur:
                    yield cur
                cur = {"name": row[1], "hdr": None, "rows": []}
            elif cur is not None:
                if cur["hdr"] is None:
                    cur["hdr"] = row
                else:
                    cur["rows"].append(row)
    if cur:
        yield cur


def main():
    path, wanted = sys.argv[1], sys.argv[2:]
    best = {}
    for s in sections(path):
        h = s["hdr"]
        ie = h.index("Instructions Executed")
        tot = sum(int(r[ie]) for r in s["rows"])
        key = s["name"].split(">(")[0] + ">"
        if wanted and not any(w in key for w in wanted):
            continue
        if key not in best or tot > best[key][0]:
            best[key] = (tot, s)
    for key, (tot, s) in sorted(best.items()):
        h = s["hdr"]
        ie, it, ss = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
        print("== %s: largest captured launch, %d warp instructions, %d SASS instructions" % (key, tot, len(s["rows"])))
        blocks = []
        for r in s["rows"]:
            c = int(r[ie])
            if blocks and blocks[-1][0] == c:
                blocks[-1][1].append(r)
            else:
                blocks.append([c, [r]])
        print("   share  executions  instr  lanes  samples  mix (first opcodes)")
        for c, rs in blocks:
            if c * len(rs) < 0.02 * tot:
                continue
            lanes = sum(int(r[it]) for r in rs) / max(1, sum(int(r[ie]) for r in rs))
            ops = {}
            for r in rs:
                op = r[1].strip().split()[0 if not r[1].strip().startswith("@") else 1].split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            mix = " ".join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
            print("  %5.1f%%  %10d  %5d  %5.1f  %7d  %s" % (100.0 * c * len(rs) / tot, c, len(rs), lanes, sum(int(r[ss]) for r in rs), mix))
        print()


if __name__ == "__main__":
    main()
