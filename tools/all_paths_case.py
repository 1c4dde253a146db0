"""Small graphs through every kernel path (spans, paired/ordinary batches, hub segments,
heavy and light chains, sequential-regime columns, fused and cascade L1, build_P)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clane_b200 import similarity, synth
from clane_b200.embedder import Embedder
from clane_b200.graph import Graph

rng = np.random.default_rng(3)
for n, d, hub in [(3000, 128, 1500), (2000, 100, 1100), (1500, 20, 300), (800, 1433, 0)]:
    src, dst = synth.make_edges(n, n * 5, "powerlaw", rng)
    if hub:
        src = np.concatenate([src, np.full(hub, 7), np.full(300, 9)])
        dst = np.concatenate([dst, rng.permutation(n)[:hub], rng.permutation(n)[:300]])
    X = rng.standard_normal((n, d), dtype=np.float32)
    g = Graph.from_arrays(n, src, dst, X)
    e = Embedder(g, similarity.CosineSimilarity(), device=torch.device("cuda"), gamma=0.76, tolerence=2)
    e.verbose = False
    e.propagate(max_sweeps=3)
    torch.cuda.synchronize()
    print("ok", n, d, "hub rows", g._device_state().plan.n_hub_rows, "sweeps", e.sweeps_per_call)
