"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference with the one-line numpy shim
SURVEY.md section 8(c) describes (``numpy.Inf`` was removed in numpy 2;
/root/reference/clane/embedder.py:1 imports it) and runs under
``torch.set_num_threads(1)`` (the oracle conditions of SURVEY.md section 7.1).

Two kinds of fixtures are written:

* ``ref_<case>.npz``  -- inputs and outputs of ``Graph`` / ``Graph.build_P`` /
  ``Embedder.iterate`` of the reference itself: coalesced indices, first-call P values,
  every per-sweep L1 amount (captured as the exact fp32 tensors the reference prints at
  embedder.py:104), per-call sweep counts, outer amounts and the final Z.
* ``prim_torch.npz`` -- outputs of the third-party primitives the reference calls
  (torch.sum, softmax, mm, bmm on torch-CPU/oneMKL) at sizes the reference itself is too
  slow to reach; inputs are regenerated from seeds by the tests.

It also writes tests/data_root/{V,E} (the Zachary karate-club toy graph the reference's
own tests use, tests/test_graph.py:13) and tests/config.yaml so config 1 of BASELINE.json
can be run by the new package without /root/reference.
"""
from __future__ import annotations

import io
import sys
import tempfile
from contextlib import redirect_stdout
from pathlib import Path

import numpy

numpy.Inf = numpy.inf  # shim, see module docstring
import numpy as np
import torch

REF = Path("/root/reference")
sys.path.insert(0, str(REF))
torch.set_num_threads(1)

import clane.embedder as ref_embedder  # noqa: E402
from clane.embedder import Embedder  # noqa: E402
from clane.graph import Graph  # noqa: E402
from clane.similarity import CosineSimilarity  # noqa: E402

HERE = Path(__file__).resolve().parent
ref_embedder.tqdm = lambda x: x  # silence the per-vertex progress bar


def run_reference(data_root: Path, d: int, gamma: float, tol: int, seed: int | None):
    """Run Graph + build_P + Embedder.iterate of the reference; capture everything."""
    if seed is not None:
        torch.manual_seed(seed)
    g = Graph(data_root=data_root, embedding_dim=d)
    X = g.X.clone()
    A = g.A
    sim = CosineSimilarity()
    P0 = g.build_P(sim)
    scores0 = sim(*[t.squeeze(0) for t in g.Z[A.indices()].split(1)])
    captured = []

    def fake_print(*args, **kwargs):
        captured.append((args[0].clone(), int(args[1])))

    ref_embedder.print = fake_print
    emb = Embedder(graph=g, similarity_measure=sim, device=torch.device("cpu"), gamma=gamma,
                   tolerence=tol, save_history=True)
    outer_amounts = []
    # replicate iterate() bookkeeping only to capture the outer amounts: run iterate() itself,
    # then recompute the outer amounts from the history with the same torch ops.
    emb.iterate()
    del ref_embedder.print
    sweeps_per_call = [len(h) for h in emb.history["Z"]]
    prev = X
    for h in emb.history["Z"]:
        outer_amounts.append((h[-1].clone() - prev).abs().sum())
        prev = h[-1]
    amounts = torch.stack([a for a, _ in captured]).numpy()
    counters = np.array([c for _, c in captured], np.int64)
    assert sum(sweeps_per_call) == len(amounts)
    raw_src = np.array([e.src.idx for e in g.E], np.int64)
    raw_dst = np.array([e.dst.idx for e in g.E], np.int64)
    nbrs = [g.get_nbrs(i).numpy() for i in range(len(g))]
    return dict(
        X=X.numpy(), n=np.int64(len(g)), d=np.int64(d), gamma=np.float64(gamma), tol=np.int64(tol),
        raw_src=raw_src, raw_dst=raw_dst,
        A_indices=A.indices().numpy(), P0_indices=P0.indices().numpy(), P0_values=P0.values().numpy(),
        scores0=scores0.numpy(),
        nbr_ptr=np.cumsum([0] + [len(x) for x in nbrs]).astype(np.int64),
        nbr_idx=np.concatenate(nbrs).astype(np.int64) if nbrs else np.zeros(0, np.int64),
        amounts=amounts, counters=counters, sweeps_per_call=np.array(sweeps_per_call, np.int64),
        outer_amounts=torch.stack(outer_amounts).numpy(),
        Z_final=g.Z.numpy(), Z_after_first_call=emb.history["Z"][0][-1].numpy(),
        Z_after_first_sweep=emb.history["Z"][0][0].numpy(),
    )


def write_graph(root: Path, ids, edges, X=None):
    root.mkdir(parents=True, exist_ok=True)
    (root / "V").write_text("\n".join(ids))
    (root / "E").write_text("\n".join(f"{a}\t{b}" for a, b in edges))
    if X is not None:
        np.save(root / "C.npy", X)


def case_random(name, n, e_raw, d, gamma, tol, seed, features="normal", hub=None, dups=0, loops=0,
                weird_ids=False):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, e_raw)
    dst = rng.integers(0, n, e_raw)
    if hub is not None:
        k = hub
        src = np.concatenate([src, np.full(k, 3)])
        dst = np.concatenate([dst, rng.permutation(n)[:k]])
    if dups:
        sel = rng.integers(0, len(src), dups)
        src = np.concatenate([src, src[sel]])
        dst = np.concatenate([dst, dst[sel]])
    if loops:
        l = rng.integers(0, n, loops)
        src = np.concatenate([src, l])
        dst = np.concatenate([dst, l])
    order = rng.permutation(len(src))
    src, dst = src[order], dst[order]
    if features == "normal":
        X = rng.standard_normal((n, d)).astype(np.float32)
    elif features == "bow":
        X = (rng.random((n, d)) < 18.0 / d).astype(np.float32)
    elif features == "tfidf":
        X = ((rng.random((n, d)) < 0.1) * rng.random((n, d)) * 0.2).astype(np.float32)
    ids = [f"v{(7 * i + 3) % n:03d}x" if weird_ids else str(i) for i in range(n)]
    if weird_ids:
        assert len(set(ids)) == n
    with tempfile.TemporaryDirectory() as tmp:
        root = Path(tmp)
        write_graph(root, ids, [(ids[a], ids[b]) for a, b in zip(src, dst)], X)
        out = run_reference(root, d, gamma, tol, None)
    out["ids"] = np.array(ids)
    np.savez_compressed(HERE / f"ref_{name}.npz", **out)
    print(name, "sweeps/call", out["sweeps_per_call"].tolist(), "E", out["A_indices"].shape[1],
          "raw", len(src), "maxdeg", int(np.diff(out["nbr_ptr"]).max()))


def case_toy(name, d, gamma, tol, seed):
    root = REF / "tests" / "data_root"
    out = run_reference(root, d, gamma, tol, seed)
    out["seed"] = np.int64(seed)
    np.savez_compressed(HERE / f"ref_{name}.npz", **out)
    print(name, "sweeps/call", out["sweeps_per_call"].tolist())


def prim_fixtures():
    """Third-party primitives at sizes the reference cannot reach (inputs rebuilt from seeds)."""
    out = {}
    # torch.sum cascade (similarity.py:37, embedder.py:94)
    sum_n = [1, 3, 7, 8, 31, 32, 33, 100, 511, 512, 513, 4096, 16384, 16385, 65536, 100000, 524288,
             524289 + 77, 1 << 20, (1 << 21) + 12345, 5_000_011, 21_675_904, 40_000_003]
    res = []
    for i, n in enumerate(sum_n):
        x = np.abs(np.random.default_rng(1000 + i).standard_normal(n).astype(np.float32))
        res.append(torch.from_numpy(x).sum().item())
    out["sum_n"] = np.array(sum_n, np.int64)
    out["sum_out"] = np.array(res, np.float32)
    # softmax rows (graph.py:123)
    sm_k = [1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 48, 100, 255, 1000, 8234]
    for i, k in enumerate(sm_k):
        for j, scale in enumerate((0.01, 1.0, 30.0)):
            s = (np.random.default_rng(2000 + 10 * i + j).standard_normal(k) * scale).astype(np.float32)
            out[f"softmax_{k}_{j}"] = torch.from_numpy(s).softmax(0).numpy()
    out["softmax_k"] = np.array(sm_k, np.int64)
    # row update  w[1,k] @ Z[k,d]   (embedder.py:92)
    mm_cases = [(1, 2), (3, 8), (7, 16), (8, 16), (9, 17), (8, 2), (16, 20), (23, 24), (40, 100), (8, 128),
                (15, 128), (64, 128), (300, 128), (2000, 128), (20000, 128), (5, 500), (33, 500), (12, 1433),
                (100, 1433), (9, 15), (50, 31), (50, 33)]
    for i, (k, d) in enumerate(mm_cases):
        rng = np.random.default_rng(3000 + i)
        w = rng.random(k).astype(np.float32)
        w /= w.sum()
        Z = rng.standard_normal((k, d)).astype(np.float32)
        out[f"mm_{k}_{d}"] = torch.from_numpy(w).view(1, -1).mm(torch.from_numpy(Z)).numpy()[0]
    out["mm_cases"] = np.array(mm_cases, np.int64)
    # batched dot (similarity.py:35-37) at d on both sides of the MKL switch
    dot_d = [2, 8, 100, 128, 399, 400, 500, 1433]
    for i, d in enumerate(dot_d):
        rng = np.random.default_rng(4000 + i)
        a = rng.standard_normal((64, d)).astype(np.float32)
        b = rng.standard_normal((64, d)).astype(np.float32)
        out[f"dot_{d}"] = torch.from_numpy(a).unsqueeze(1).matmul(torch.from_numpy(b).unsqueeze(-1)).view(-1).numpy()
    out["dot_d"] = np.array(dot_d, np.int64)
    np.savez_compressed(HERE / "prim_torch.npz", **out)
    print("prim fixtures written")


def broadcast_fixtures():
    """CosineSimilarity of the reference on a single row against a batch (similarity.py:35-37 broadcasts)."""
    out = {}
    cases = [(5, 8), (64, 128), (20, 500)]
    sim = CosineSimilarity()
    for i, (e, d) in enumerate(cases):
        rng = np.random.default_rng(5000 + i)
        a = rng.standard_normal((1, d)).astype(np.float32)
        b = rng.standard_normal((e, d)).astype(np.float32)
        out[f"ab_{e}_{d}"] = sim(torch.from_numpy(a), torch.from_numpy(b)).numpy()
        out[f"ba_{e}_{d}"] = sim(torch.from_numpy(b), torch.from_numpy(a)).numpy()
    out["cases"] = np.array(cases, np.int64)
    np.savez_compressed(HERE / "prim_cosine_broadcast.npz", **out)
    print("broadcast fixtures written")


def main():
    if sys.argv[1:] == ["broadcast"]:
        return broadcast_fixtures()
    broadcast_fixtures()
    # the toy graph + config the reference's own tests use (tests/test_graph.py:13, tests/config.yaml)
    root = HERE.parent / "data_root"
    root.mkdir(exist_ok=True)
    for f in ("V", "E"):
        (root / f).write_bytes((REF / "tests" / "data_root" / f).read_bytes())
    (HERE.parent / "config.yaml").write_text(
        'graph:\n  embedding_dim: 2\n\nsimilarity:\n  method: "CosineSimilarity"\n  kwargs:\n    foo: "bar"\n\n'
        "embedder:\n  gamma: 0.76\n  tolerence: 10\n")
    prim_fixtures()
    case_toy("toy_d2", 2, 0.76, 10, 0)          # BASELINE.json config 1
    case_toy("toy_d16", 16, 0.74, 10, 0)        # tests/test_embedder.py:11-29
    case_random("cyclic100_d8", 100, 390, 8, 0.76, 3, seed=0, dups=4)
    case_random("hub60_d20", 60, 500, 20, 0.76, 3, seed=1, hub=51, dups=20, loops=6, weird_ids=True)
    case_random("bow40_d1433", 40, 290, 1433, 0.76, 2, seed=2, features="bow")
    case_random("tfidf30_d500", 30, 200, 500, 0.76, 2, seed=3, features="tfidf", hub=24)
    case_random("n50_d100", 50, 400, 100, 0.76, 2, seed=4, hub=40)
    case_random("n120_d128", 120, 1100, 128, 0.5, 2, seed=5, hub=100, dups=10, loops=3)


if __name__ == "__main__":
    main()
