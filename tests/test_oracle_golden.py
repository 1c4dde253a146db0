"""The oracle (oracle/clane_oracle.c) against the golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  Bit-exact everywhere: this is what pins the oracle."""
import glob
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
REF_CASES = sorted(p.stem for p in GOLD.glob("ref_*.npz"))


@pytest.fixture(scope="module")
def prim():
    return np.load(GOLD / "prim_torch.npz")


def test_golden_files_present():
    assert len(REF_CASES) >= 8 and (GOLD / "prim_torch.npz").exists()


@pytest.mark.parametrize("threads", [1, 4])
def test_cascade_sum_matches_torch_sum(prim, threads):
    O.set_threads(threads)
    for i, n in enumerate(prim["sum_n"]):
        x = np.abs(np.random.default_rng(1000 + i).standard_normal(int(n)).astype(np.float32))
        assert O.aten_sum(x) == prim["sum_out"][i], f"n={n}"


def test_softmax_rows_match_torch(prim):
    for i, k in enumerate(prim["softmax_k"]):
        for j, scale in enumerate((0.01, 1.0, 30.0)):
            s = (np.random.default_rng(2000 + 10 * i + j).standard_normal(int(k)) * scale).astype(np.float32)
            w = O.softmax_rows(s, np.array([0, k]))
            assert np.array_equal(w, prim[f"softmax_{k}_{j}"]), f"k={k} scale={scale}"


def test_row_update_matches_mkl_sgemm(prim):
    for i, (k, d) in enumerate(prim["mm_cases"]):
        rng = np.random.default_rng(3000 + i)
        w = rng.random(k).astype(np.float32)
        w /= w.sum()
        Z = rng.standard_normal((k, d)).astype(np.float32)
        Zc = np.concatenate([np.zeros((1, d), np.float32), Z])   # node 0 -> nodes 1..k, x = 0, gamma = 1
        rowptr = np.zeros(k + 2, np.int64)
        rowptr[1:] = k
        Zn = O.sweep(np.zeros_like(Zc), Zc, rowptr, np.arange(1, k + 1, dtype=np.int32), w, 1.0)
        assert np.array_equal(Zn[0], prim[f"mm_{k}_{d}"]), f"k={k} d={d}"
        assert np.array_equal(Zn[1:], Zc[1:])                      # sinks keep their value


def test_batched_dot_matches_torch(prim):
    for i, d in enumerate(prim["dot_d"]):
        rng = np.random.default_rng(4000 + i)
        a = rng.standard_normal((64, d)).astype(np.float32)
        b = rng.standard_normal((64, d)).astype(np.float32)
        rowptr = np.concatenate([np.arange(65), np.full(64, 64)]).astype(np.int64)
        dots, _, _ = O.scores_raw(np.concatenate([a, b]), rowptr, np.arange(64, 128, dtype=np.int32))
        assert np.array_equal(dots, prim[f"dot_{d}"]), f"d={d}"


@pytest.mark.parametrize("case", REF_CASES)
def test_oracle_reproduces_reference(case):
    G = np.load(GOLD / f"{case}.npz")
    n = int(G["n"])
    O.set_threads(2)
    rowptr, col = O.csr_from_edges(G["raw_src"], G["raw_dst"], n)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    assert np.array_equal(rows, G["A_indices"][0]) and np.array_equal(col, G["A_indices"][1])   # graph.py:104-110
    assert np.array_equal(rowptr, G["nbr_ptr"]) and np.array_equal(col, G["nbr_idx"])           # get_nbrs
    dots, s1, s2 = O.scores_raw(G["X"], rowptr, col)
    c = np.float32(np.sqrt(s1, dtype=np.float32) * np.sqrt(s2, dtype=np.float32))
    assert np.array_equal((dots / c).astype(np.float32), G["scores0"])                           # similarity.py:37
    assert np.array_equal(O.build_p(G["X"], rowptr, col), G["P0_values"])                        # graph.py:118-128
    gamma, tol = float(G["gamma"]), int(G["tol"])
    w = O.build_p(G["X"], rowptr, col)
    assert np.array_equal(O.sweep(G["X"], G["X"], rowptr, col, w, gamma), G["Z_after_first_sweep"])
    Z1, amounts, _ = O.propagate(G["X"], G["X"], rowptr, col, gamma, tol)
    k0 = int(G["sweeps_per_call"][0])
    assert len(amounts) == k0 and np.array_equal(amounts, G["amounts"][:k0])                     # embedder.py:94
    assert np.array_equal(Z1, G["Z_after_first_call"])
    Z, spc, oam = O.iterate(G["X"], rowptr, col, gamma, tol)
    assert spc.tolist() == G["sweeps_per_call"].tolist()                                         # iteration counts
    assert np.array_equal(oam, G["outer_amounts"])
    assert np.array_equal(Z, G["Z_final"])


def test_oracle_threads_do_not_change_results():
    G = np.load(GOLD / "ref_n120_d128.npz")
    n = int(G["n"])
    rowptr, col = O.csr_from_edges(G["raw_src"], G["raw_dst"], n)
    outs = []
    for t in (1, 3, 8):
        O.set_threads(t)
        outs.append(O.iterate(G["X"], rowptr, col, float(G["gamma"]), int(G["tol"])))
    for Z, spc, oam in outs[1:]:
        assert np.array_equal(Z, outs[0][0]) and spc.tolist() == outs[0][1].tolist()


def test_cosine_single_row_against_batch_matches_reference():
    """The reference's matmul broadcasts a [1, d] argument (similarity.py:35-37): dots of the expanded batch in the
    usual order, each global norm over its tensor as given.  Golden values of the real reference."""
    G = np.load(GOLD / "prim_cosine_broadcast.npz")
    for i, (e, d) in enumerate(G["cases"]):
        rng = np.random.default_rng(5000 + i)
        a = rng.standard_normal((1, d)).astype(np.float32)
        b = rng.standard_normal((e, d)).astype(np.float32)
        rowptr = np.concatenate([np.arange(e + 1), np.full(e, e)]).astype(np.int64)
        dots, _, _ = O.scores_raw(np.concatenate([np.repeat(a, e, 0), b]), rowptr, np.arange(e, 2 * e, dtype=np.int32))
        c = np.float32(np.sqrt(O.aten_sum(a * a), dtype=np.float32) * np.sqrt(O.aten_sum(b * b), dtype=np.float32))
        want = (dots / c).astype(np.float32)
        assert np.array_equal(want, G[f"ab_{e}_{d}"]) and np.array_equal(want, G[f"ba_{e}_{d}"])
