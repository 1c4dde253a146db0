"""Synthetic graphs of the shapes BASELINE.json names (configs 2-5).

One ``numpy.random.default_rng(seed)`` per config; draw order: u, v (oversampled x1.3),
perm_s, perm_d, the subset permutation, then X.  Edges are returned in file (unsorted) order.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: (nodes, edges, dim, edge model, feature model)
    "cora": (2708, 5429, 1433, "uniform", "bow"),
    "pubmed": (19717, 44338, 500, "uniform", "tfidf"),
    "arxiv": (169343, 1166243, 128, "powerlaw", "normal"),
    "products": (2449029, 61859140, 100, "powerlaw", "normal"),
}


def make_edges(n: int, e: int, model: str, rng) -> tuple[np.ndarray, np.ndarray]:
    m = int(e * 1.3)
    u, v = rng.random(m), rng.random(m)
    perm_s, perm_d = rng.permutation(n), rng.permutation(n)
    if model == "powerlaw":   # density ~ rank^-0.6
        src = perm_s[np.floor(n * u ** 2.5).astype(np.int64)]
        dst = perm_d[np.floor(n * v ** 2.5).astype(np.int64)]
    else:
        src = perm_s[np.floor(n * u).astype(np.int64)]
        dst = perm_d[np.floor(n * v).astype(np.int64)]
    key = src * n + dst
    key = np.unique(key[src != dst])
    if len(key) < e:
        raise ValueError(f"oversampling produced only {len(key)} unique edges, need {e}")
    key = key[rng.permutation(len(key))[:e]]
    return key // n, key % n


def make_features(n: int, d: int, model: str, rng) -> np.ndarray:
    if model == "bow":      # ~18 words per document, 0/1
        return (rng.random((n, d)) < 18.0 / d).astype(np.float32)
    if model == "tfidf":    # Bernoulli(0.1) mask x Uniform(0, 0.2)
        return ((rng.random((n, d)) < 0.1) * rng.random((n, d)) * 0.2).astype(np.float32)
    return rng.standard_normal((n, d), dtype=np.float32)


def make_graph(shape: str, seed: int = 0, scale: float = 1.0):
    """(n, src, dst, X) for a named shape; ``scale`` shrinks nodes and edges proportionally."""
    n, e, d, emodel, fmodel = SHAPES[shape]
    n, e = max(int(n * scale), 8), max(int(e * scale), 8)
    rng = np.random.default_rng(seed)
    src, dst = make_edges(n, e, emodel, rng)
    X = make_features(n, d, fmodel, rng)
    return n, src, dst, X
