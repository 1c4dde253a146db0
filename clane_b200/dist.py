"""Row-partitioned multi-GPU sweep: one process per GPU, NCCL over NVLink.

Nodes are partitioned by contiguous id range; every rank keeps the full CSR and a full replica
of Z (out-neighbours span all ranks), sweeps its own rows, and the ranks exchange their Znext
slices with one all-gather per sweep.  The three global scalars stay bit-identical on every rank:

  * L1 change per sweep (embedder.py:94): after the all-gather each rank reduces a disjoint range
    of the ATen cascade's level-1 nodes over the full [N*d] array (clane_l1_partial), the ranks
    all-reduce(SUM) the node slots -- every slot is written by exactly one rank and is +0
    elsewhere, so the sum is exact -- and each rank finishes levels 2-3 and the patience state
    machine itself (clane_l1_finish).  Patience is replicated, never broadcast.
  * the two Frobenius norms of build_P (similarity.py:37) are computed redundantly by every rank
    over all E*d gathered elements (each rank holds the full Z and CSR); dots and softmax only for
    the rank's own rows.

The helpers at the top are pure host logic (tested under gloo on CPU); ShardedSweeper needs CUDA.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

ROW_ALIGN = 8   # slices are cut at multiples of the sweep's row-group size


def rows_per_rank(n: int, world: int) -> int:
    """Equal slice length (a multiple of ROW_ALIGN) such that world * length >= n."""
    per = -(-n // world)
    return -(-per // ROW_ALIGN) * ROW_ALIGN


def row_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) owned by `rank`; trailing ranks may own fewer (or no) rows."""
    per = rows_per_rank(n, world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def node_range(n_nodes: int, world: int, rank: int) -> tuple[int, int]:
    """Level-1 cascade nodes [lo, hi) reduced by `rank` (contiguous, balanced)."""
    base, rem = divmod(n_nodes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, full: torch.Tensor, group=None) -> None:
    """all-gather equal-length row slices into `full` ([world * per, ld], contiguous)."""
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)


class ShardedSweeper:
    """Sweeps of one propagate() call over this rank's rows, with the per-sweep exchange."""

    def __init__(self, graph, similarity, gamma: float, tol: int = 10, max_sweeps: int = 0):
        self.g, self.sim, self.gamma = graph, similarity, float(np.float32(gamma))
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.dev = _lib.require_cuda()
        L = _lib.lib()
        n, e, d = graph._n, graph._nnz, int(graph.X.shape[1])
        ld = int(L.clane_padded_ld(d))
        self.n, self.e, self.d, self.ld = n, e, d, ld
        self.per = rows_per_rank(n, self.world)
        self.lo, self.hi = row_range(n, self.world, self.rank)
        npad = self.per * self.world
        dev = self.dev
        self.rowptr = torch.from_numpy(graph._rowptr).to(dev)
        self.col = torch.from_numpy(graph._col if e else np.zeros(1, np.int32)).to(dev)
        self.erow = torch.zeros(max(e, 1), dtype=torch.int32, device=dev)
        self.X = torch.zeros([npad, ld], dtype=torch.float32, device=dev)
        self.X[:n, :d] = graph.X.to(dev)
        self.Z = [self.X.clone(), self.X.clone()]
        self.cur = 0
        self.w = torch.zeros(max(e, 1), dtype=torch.float32, device=dev)
        self.norms2 = torch.zeros(2, dtype=torch.float32, device=dev)
        self.amount = torch.zeros(1, dtype=torch.float32, device=dev)
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)
        self.state_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.log_cap = 1 << 16
        self.log = torch.zeros(self.log_cap, dtype=torch.float32, device=dev)
        self.plan = _lib.Plan(n, e, d, graph._rowptr, self.lo, self.hi, 0)
        nodes = ctypes.c_int64()
        _lib.check(L.clane_cascade_shape(n * d, ctypes.byref(nodes), None))
        self.n1 = int(nodes.value)
        self.nlo, self.nhi = node_range(self.n1, self.world, self.rank)
        self.p1 = torch.zeros((self.n1 + 2) * 32, dtype=torch.float32, device=dev)
        s = _lib.stream_handle()
        _lib.check(L.clane_edge_rows(self.rowptr.data_ptr(), n, e, self.erow.data_ptr(), s))
        self.tol, self.max_sweeps = tol, max_sweeps
        self.build_p()
        _lib.check(L.clane_patience_reset(self.state.data_ptr(), tol, max_sweeps, s))

    def build_p(self) -> None:
        L = _lib.lib()
        _lib.check(L.clane_build_p_cosine(self.plan.handle, self.Z[self.cur].data_ptr(), self.rowptr.data_ptr(),
                                          self.erow.data_ptr(), self.col.data_ptr(), self.w.data_ptr(),
                                          self.norms2.data_ptr(), _lib.stream_handle()), "clane_build_p_cosine")

    def sweep(self, with_l1: bool = True) -> int:
        """One sweep: own rows, all-gather of the new slices, exact L1 + patience.  Returns the
        number of kernels this rank launched."""
        L = _lib.lib()
        s = _lib.stream_handle()
        zc, zn = self.Z[self.cur], self.Z[self.cur ^ 1]
        _lib.check(L.clane_sweep(self.plan.handle, self.X.data_ptr(), zc.data_ptr(), zn.data_ptr(), self.rowptr.data_ptr(),
                                 self.col.data_ptr(), self.w.data_ptr(), ctypes.c_float(self.gamma), 0, 0, 0, 0, s),
                   "clane_sweep")
        launches = 1 + (1 if self.plan.n_hub_rows else 0)
        lo = self.rank * self.per
        gather_rows(zn[lo:lo + self.per], zn)
        if with_l1:
            self.p1.zero_()
            _lib.check(L.clane_l1_partial(self.plan.handle, zn.data_ptr(), zc.data_ptr(), self.nlo, self.nhi,
                                          self.p1.data_ptr(), s), "clane_l1_partial")
            dist.all_reduce(self.p1, op=dist.ReduceOp.SUM)
            _lib.check(L.clane_l1_finish(self.plan.handle, zn.data_ptr(), zc.data_ptr(), self.p1.data_ptr(),
                                         self.amount.data_ptr(), 0, 0, 0, s), "clane_l1_finish")
            launches += 3
        self.cur ^= 1
        return launches

    def last_amount(self) -> float:
        return float(self.amount.cpu()[0])

    def Z_host(self) -> torch.Tensor:
        return self.Z[self.cur][:self.n, :self.d].cpu()
