import os, sys, ctypes
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from clane_b200 import similarity, synth, _lib
from clane_b200.embedder import Embedder
from clane_b200.graph import Graph
shape, scale = sys.argv[1], float(sys.argv[2])
n, src, dst, X = synth.make_graph(shape, seed=0, scale=scale)
g = Graph.from_arrays(n, src, dst, X)
e = Embedder(g, similarity.CosineSimilarity(), device=torch.device("cuda"), gamma=0.76, tolerence=10)
e.verbose = False
e.propagate(max_sweeps=2)
torch.cuda.synchronize()
print("deg max", np.diff(g._rowptr).max(), "hub rows", g._device_state().plan.n_hub_rows)
