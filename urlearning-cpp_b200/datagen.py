"""Seeded synthetic inputs for the BASELINE.json configs.

The reference's own generators (score/generate_chain.cpp:81-86, score/generate_sample_data.cpp:70) are
time-seeded toys with N=1000 hard-coded; these are deterministic numpy generators of the named shapes.
"""
from __future__ import annotations

import numpy as np


def discrete_bn(p: int = 60, n: int = 1_000_000, seed: int = 4, window: int = 4, max_indegree: int = 3,
                arities=(2, 3, 4), alpha: float = 0.5, sample_seed: int | None = None):
    """Config 4: discrete Bayesian network, forward sampled.

    Variable i draws up to ``max_indegree`` parents from the ``window`` preceding variables, so the undirected
    generating graph (the skeleton handed to ``score -k``, standing in for MMPC) has degree <= 2*window and the
    2-hop neighbourhood of any variable lies inside [i-2*window, i+2*window].  CPT rows ~ Dirichlet(alpha).

    ``sample_seed``: draw the CPTs and the rows from an independent generator, keeping the arities and the DAG of
    ``seed`` (replicas of one network structure with their own data: the blocks of bench.py's weak-scaling data set).

    Returns (codes uint8 [p, n], card int32 [p], edges list[int] skeleton masks, parents list[list[int]]).
    """
    rng = np.random.default_rng(seed)
    card = rng.choice(np.asarray(arities), size=p).astype(np.int32)
    parents = []
    for i in range(p):
        lo = max(0, i - window)
        cands = np.arange(lo, i)
        k = min(max_indegree, len(cands))
        parents.append(sorted(rng.choice(cands, size=k, replace=False).tolist()) if k else [])
    if sample_seed is not None:
        rng = np.random.default_rng(sample_seed)
    codes = np.zeros((p, n), dtype=np.uint8)
    for i in range(p):
        pa = parents[i]
        ncfg = int(np.prod([card[j] for j in pa])) if pa else 1
        cpt = rng.dirichlet(np.full(card[i], alpha), size=ncfg)  # [ncfg, r_i]
        cfg = np.zeros(n, dtype=np.int64)
        mult = 1
        for j in pa:
            cfg += codes[j].astype(np.int64) * mult
            mult *= int(card[j])
        cum = np.cumsum(cpt, axis=1)
        u = rng.random(n)
        val = (u[:, None] > cum[cfg]).sum(axis=1)
        codes[i] = np.minimum(val, card[i] - 1).astype(np.uint8)
    # codes are value indices in first-appearance order in the reference (variable.h:43-48); relabel so that the
    # arrays equal what the reference's reader would produce from a CSV of these values
    for i in range(p):
        _, first = np.unique(codes[i], return_index=True)
        order = np.argsort(first)
        present = np.unique(codes[i])
        remap = np.zeros(256, dtype=np.uint8)
        remap[present[order]] = np.arange(len(present), dtype=np.uint8)
        codes[i] = remap[codes[i]]
        card[i] = len(present)
    edges = [0] * p
    for i in range(p):
        for j in parents[i]:
            edges[i] |= 1 << j
            edges[j] |= 1 << i
    return codes, card, edges, parents


def linear_gaussian_sem(p: int = 30, n: int = 100_000, seed: int = 3, mean_indegree: float = 2.0):
    """Config 3: linear-Gaussian SEM over a random DAG (random topological order, each node picks parents among
    its predecessors with probability giving ``mean_indegree`` on average), weights +-U[0.5,1.5], unit noise.

    Returns (x float64 [p, n], dag parents list[list[int]]).
    """
    rng = np.random.default_rng(seed)
    order = rng.permutation(p)
    parents = [[] for _ in range(p)]
    x = np.zeros((p, n), dtype=np.float64)
    for pos, v in enumerate(order):
        if pos:
            prob = min(1.0, mean_indegree / pos) if pos > mean_indegree else 1.0
            pick = rng.random(pos) < prob
            parents[v] = sorted(int(u) for u in order[:pos][pick])
        val = rng.standard_normal(n)
        for u in parents[v]:
            w = rng.uniform(0.5, 1.5) * rng.choice([-1.0, 1.0])
            val += w * x[u]
        x[v] = val
    return x, parents


def write_csv(path: str, cols: np.ndarray, header=None, fmt="%d"):
    """cols [p, n] -> CSV with one record per line (the reference's input format)."""
    with open(path, "w") as f:
        if header:
            f.write(",".join(header) + "\n")
        np.savetxt(f, cols.T, fmt=fmt, delimiter=",")


def write_skeleton_matrix(path: str, edges, p: int):
    """0/1 matrix without blank lines (skeleton.cpp:84-99 counts every line as a row)."""
    with open(path, "w") as f:
        for i in range(p):
            f.write(",".join("1" if (edges[i] >> j) & 1 else "0" for j in range(p)) + "\n")
