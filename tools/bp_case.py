"""build_P a few times on a named shape (for an ncu launch list).  python tools/bp_case.py [workload]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clane_b200 import similarity, synth  # noqa: E402
from clane_b200.graph import Graph  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
n, src, dst, X = synth.make_graph(name, seed=0)
g = Graph.from_arrays(n, src, dst, X)
for _ in range(3):
    g._build_P_device(similarity.CosineSimilarity())
torch.cuda.synchronize()
print("ok")
