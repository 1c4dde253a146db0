"""Debug aid: one sweep on the arxiv-shape graph vs the oracle; report which rows differ."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clane_b200 import similarity, synth
from clane_b200.embedder import Embedder
from clane_b200.graph import Graph
from oracle import oracle as O

n, src, dst, X = synth.make_graph("arxiv", seed=0)
g = Graph.from_arrays(n, src, dst, X)
e = Embedder(g, similarity.CosineSimilarity(), device=torch.device("cuda"), gamma=0.76, tolerence=10)
e.verbose = False
e.propagate(max_sweeps=1)
O.set_threads(O.max_threads())
rowptr, col = O.csr_from_edges(src, dst, n)
Zo, amounts, w = O.propagate(X, X, rowptr, col, 0.76, 10, max_sweeps=1)
Z = g.Z.numpy()
bad = np.nonzero((Z != Zo).any(1))[0]
deg = np.diff(rowptr)
print("amount gpu", e.amounts_per_call[0], "oracle", amounts, "bad rows", len(bad))
if len(bad):
    print("first bad rows", bad[:20], "deg", deg[bad[:20]], "row%8", bad[:20] % 8)
    print("bad deg histogram", np.bincount(np.minimum(deg[bad], 200))[:140].nonzero()[0][:40])
    r = bad[0]
    print("row", r, "maxabs", np.abs(Z[r] - Zo[r]).max(), "cols differing", np.nonzero(Z[r] != Zo[r])[0][:16])
    grp = r // 8 * 8
    print("group degrees", deg[grp:grp + 8])
