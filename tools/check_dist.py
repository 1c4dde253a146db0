"""torchrun check (N >= 2 GPUs): the row-partitioned sweeps equal the single-GPU ones bit for bit."""
import os, sys
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clane_b200 import similarity, synth, dist as cdist
from clane_b200.embedder import Embedder
from clane_b200.graph import Graph

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok = True
for shape, scale in (("arxiv", 0.3), ("products", 0.01), ("pubmed", 1.0), ("cora", 1.0)):
    n, src, dst, X = synth.make_graph(shape, seed=0, scale=scale)
    g = Graph.from_arrays(n, src, dst, X)
    sim = similarity.CosineSimilarity()
    sw = cdist.ShardedSweeper(g, sim, 0.76, exchange=os.environ.get("CLANE_EXCHANGE", "auto"))
    if rank == 0:
        print(f"exchange={sw.exchange} aligned={sw.aligned} rows/rank={sw.per}" +
              (f" (symmetric memory unavailable: {sw.symm_error})" if hasattr(sw, "symm_error") else ""), flush=True)
    amounts = []
    for _ in range(4):
        sw.sweep(True)
        amounts.append(sw.last_amount())
    Zs = sw.Z_host().numpy()
    g1 = Graph.from_arrays(n, src, dst, X)
    e = Embedder(g1, sim, device=torch.device("cuda", lr), gamma=0.76, tolerence=10)
    e.verbose = False
    e.propagate(max_sweeps=4)
    same = np.array_equal(Zs, g1.Z.numpy()) and np.array_equal(np.float32(amounts), e.amounts_per_call[0])
    # the pipelined call (L1 of sweep t beside sweep t + 1), twice in a row from the same sweeper
    sw2 = cdist.ShardedSweeper(g, sim, 0.76, exchange=os.environ.get("CLANE_EXCHANGE", "auto"))
    am2 = sw2.sweeps(3) + sw2.sweeps(1)
    same = same and np.array_equal(sw2.Z_host().numpy(), Zs) and np.array_equal(np.float32(am2), np.float32(amounts))
    if rank == 0:
        print(f"  pipelined={sw2.pipelined}", flush=True)
    del sw2
    t = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{shape} x{scale}: n={n} e={g._nnz} world={world} sharded == single-GPU: {bool(t.item())}", flush=True)
    ok &= bool(t.item())
# the whole iterate() loop, sharded vs single GPU: same sweep counts per call, same final embeddings
n, src, dst, X = synth.make_graph("arxiv", seed=1, scale=0.05)
g = Graph.from_arrays(n, src, dst, X)
sw = cdist.ShardedSweeper(g, similarity.CosineSimilarity(), 0.76, tol=3, exchange=os.environ.get("CLANE_EXCHANGE", "auto"))
spc, outer_amounts = sw.iterate()
g1 = Graph.from_arrays(n, src, dst, X)
e = Embedder(g1, similarity.CosineSimilarity(), device=torch.device("cuda", lr), gamma=0.76, tolerence=3)
e.verbose = False
e.iterate()
same = spc == e.sweeps_per_call and np.array_equal(sw.Z_host().numpy(), g1.Z.numpy())
t = torch.tensor([int(same)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"iterate(): outer={len(spc)} sweeps={sum(spc)} world={world} sharded == single-GPU: {bool(t.item())}", flush=True)
ok &= bool(t.item())
dist.destroy_process_group()
sys.exit(0 if ok else 1)
