"""ctypes front-end of the CPU oracle (oracle/clane_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  Nothing under ``clane_b200/``
imports this module; the product path fails loudly without its CUDA library instead.

The functions mirror the reference call sites they restate:
  csr_from_edges  -> Graph.A            (/root/reference/clane/graph.py:104-110)
  build_p         -> Graph.build_P      (/root/reference/clane/graph.py:118-128,
                                          /root/reference/clane/similarity.py:26-37)
  sweep           -> Embedder.propagate (/root/reference/clane/embedder.py:84-92)
  l1_diff         -> (Z - Z_cur).abs().sum() (/root/reference/clane/embedder.py:94, :60)
  propagate / iterate -> /root/reference/clane/embedder.py:71-108 / :56-69
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libclane_oracle.so"
_lib = None

_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = _HERE / "clane_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.clane_oracle_set_threads.argtypes = [C.c_int]
        L.clane_oracle_max_threads.restype = C.c_int
        L.clane_oracle_csr_from_edges.argtypes = [_i64p, _i64p, C.c_int64, C.c_int64, _i64p, _i32p]
        L.clane_oracle_csr_from_edges.restype = C.c_int64
        L.clane_oracle_aten_sum.argtypes = [_f32p, C.c_int64]
        L.clane_oracle_aten_sum.restype = C.c_float
        L.clane_oracle_l1_diff.argtypes = [_f32p, _f32p, C.c_int64]
        L.clane_oracle_l1_diff.restype = C.c_float
        L.clane_oracle_scores_raw.argtypes = [_f32p, C.c_int64, C.c_int64, _i64p, _i32p, _f32p, _f32p, _f32p]
        L.clane_oracle_expf.argtypes = [_f32p, _f32p, C.c_int64]
        L.clane_oracle_softmax_rows.argtypes = [_f32p, C.c_int64, _i64p, _f32p]
        L.clane_oracle_build_p.argtypes = [_f32p, C.c_int64, C.c_int64, _i64p, _i32p, _f32p]
        L.clane_oracle_sweep.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int64, _i64p, _i32p, _f32p, C.c_float]
        L.clane_oracle_sweep_range.argtypes = [_f32p, _f32p, _f32p, C.c_int64, C.c_int64, C.c_int64, _i64p, _i32p,
                                               _f32p, C.c_float]
        L.clane_oracle_propagate.argtypes = [_f32p, _f32p, C.c_int64, C.c_int64, _i64p, _i32p, C.c_float,
                                             C.c_int64, C.c_int64, _f32p, C.c_int64, _f32p]
        L.clane_oracle_propagate.restype = C.c_int64
        L.clane_oracle_iterate.argtypes = [_f32p, _f32p, C.c_int64, C.c_int64, _i64p, _i32p, C.c_float,
                                           C.c_int64, C.c_int64, _i64p, _f32p, C.c_int64]
        L.clane_oracle_iterate.restype = C.c_int64
        _lib = L
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)


def set_threads(t: int) -> None:
    lib().clane_oracle_set_threads(int(t))


def max_threads() -> int:
    return int(lib().clane_oracle_max_threads())


def csr_from_edges(src, dst, n: int):
    """Sorted-unique CSR of the raw edge list: (rowptr int64[n+1], col int32[E])."""
    src, ps = _i64(src)
    dst, pd = _i64(dst)
    rowptr = np.zeros(n + 1, np.int64)
    col = np.zeros(max(len(src), 1), np.int32)
    e = lib().clane_oracle_csr_from_edges(ps, pd, len(src), n, rowptr.ctypes.data_as(_i64p),
                                          col.ctypes.data_as(_i32p))
    if e < 0:
        raise ValueError("edge endpoint out of range")
    return rowptr, col[:e].copy()


def aten_sum(x) -> np.float32:
    x, p = _f32(np.asarray(x).reshape(-1))
    return np.float32(lib().clane_oracle_aten_sum(p, x.size))


def l1_diff(a, b) -> np.float32:
    a, pa = _f32(a)
    b, pb = _f32(b)
    assert a.size == b.size
    return np.float32(lib().clane_oracle_l1_diff(pa, pb, a.size))


def expf(x):
    x, p = _f32(x)
    out = np.empty_like(x)
    lib().clane_oracle_expf(p, out.ctypes.data_as(_f32p), x.size)
    return out


def scores_raw(Z, rowptr, col):
    """(dots[E], S1, S2): per-edge sequential dots and the two cascade-ordered square sums."""
    Z, pz = _f32(Z)
    rowptr, pr = _i64(rowptr)
    col, pc = _i32(col)
    n, d = Z.shape
    dots = np.empty(len(col), np.float32)
    s1, s2 = C.c_float(), C.c_float()
    lib().clane_oracle_scores_raw(pz, n, d, pr, pc, dots.ctypes.data_as(_f32p), C.byref(s1), C.byref(s2))
    return dots, np.float32(s1.value), np.float32(s2.value)


def softmax_rows(scores, rowptr):
    scores, ps = _f32(scores)
    rowptr, pr = _i64(rowptr)
    w = np.empty_like(scores)
    lib().clane_oracle_softmax_rows(ps, len(rowptr) - 1, pr, w.ctypes.data_as(_f32p))
    return w


def build_p(Z, rowptr, col):
    Z, pz = _f32(Z)
    rowptr, pr = _i64(rowptr)
    col, pc = _i32(col)
    n, d = Z.shape
    w = np.empty(len(col), np.float32)
    lib().clane_oracle_build_p(pz, n, d, pr, pc, w.ctypes.data_as(_f32p))
    return w


def sweep(X, Zcur, rowptr, col, w, gamma: float):
    X, px = _f32(X)
    Zcur, pz = _f32(Zcur)
    rowptr, pr = _i64(rowptr)
    col, pc = _i32(col)
    w, pw = _f32(w)
    n, d = X.shape
    Zn = np.empty_like(Zcur)
    lib().clane_oracle_sweep(px, pz, Zn.ctypes.data_as(_f32p), n, d, pr, pc, pw, C.c_float(np.float32(gamma)))
    return Zn


def sweep_range(X, Zcur, Znext, lo: int, hi: int, rowptr, col, w, gamma: float):
    """Sweep rows [lo, hi) only, writing into the caller's Znext (float32, C-contiguous)."""
    X, px = _f32(X)
    Zcur, pz = _f32(Zcur)
    rowptr, pr = _i64(rowptr)
    col, pc = _i32(col)
    w, pw = _f32(w)
    assert Znext.dtype == np.float32 and Znext.flags.c_contiguous and Znext.shape == Zcur.shape
    lib().clane_oracle_sweep_range(px, pz, Znext.ctypes.data_as(_f32p), lo, hi, X.shape[1], pr, pc, pw,
                                   C.c_float(np.float32(gamma)))
    return Znext


def propagate(X, Z, rowptr, col, gamma: float, tol: int, max_sweeps: int = 0, cap: int = 100000):
    """One propagate() call.  Returns (Z_new, amounts[sweeps], w[E])."""
    X, px = _f32(X)
    Z = np.array(Z, dtype=np.float32, order="C", copy=True)
    rowptr, pr = _i64(rowptr)
    col, pc = _i32(col)
    n, d = X.shape
    amounts = np.zeros(cap, np.float32)
    w = np.empty(len(col), np.float32)
    s = lib().clane_oracle_propagate(px, Z.ctypes.data_as(_f32p), n, d, pr, pc, C.c_float(np.float32(gamma)),
                                     tol, max_sweeps, amounts.ctypes.data_as(_f32p), cap,
                                     w.ctypes.data_as(_f32p))
    return Z, amounts[:min(s, cap)].copy(), w


def iterate(X, rowptr, col, gamma: float, tol: int, Z0=None, max_outer: int = 0, cap: int = 4096):
    """Full iterate().  Returns (Z, sweeps_per_call[outer], outer_amounts[outer])."""
    X, px = _f32(X)
    Z = np.array(X if Z0 is None else Z0, dtype=np.float32, order="C", copy=True)
    rowptr, pr = _i64(rowptr)
    col, pc = _i32(col)
    n, d = X.shape
    spc = np.zeros(cap, np.int64)
    oam = np.zeros(cap, np.float32)
    o = lib().clane_oracle_iterate(px, Z.ctypes.data_as(_f32p), n, d, pr, pc, C.c_float(np.float32(gamma)),
                                   tol, max_outer, spc.ctypes.data_as(_i64p), oam.ctypes.data_as(_f32p), cap)
    return Z, spc[:o].copy(), oam[:o].copy()
