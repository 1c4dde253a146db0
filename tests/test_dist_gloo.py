"""world_size-2 gloo tests (CPU) of the host logic of the row-partitioned path (clane_b200/dist.py):
slice geometry, the all-gather of row slices, and the exact combination of disjoint cascade slots."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clane_b200 import dist as cdist


def test_partition_geometry():
    for n in (1, 7, 8, 9, 34, 2708, 169343, 2449029):
        for world in (1, 2, 3, 4, 8):
            per = cdist.rows_per_rank(n, world)
            assert per % cdist.ROW_ALIGN == 0 and per * world >= n
            covered = []
            for r in range(world):
                lo, hi = cdist.row_range(n, world, r)
                assert 0 <= lo <= hi <= n and lo % cdist.ROW_ALIGN == 0 or lo == n
                covered += list(range(lo, hi)) if n < 5000 else []
                if r:
                    assert lo == min(prev_hi, n) if hi > lo else True
                prev_hi = hi
            if n < 5000:
                assert covered == list(range(n))
            assert cdist.row_range(n, world, world - 1)[1] == n
    for nodes in (0, 1, 5, 662, 7477):
        for world in (1, 2, 4, 8):
            ranges = [cdist.node_range(nodes, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nodes
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_node_aligned_partition():
    """Cuts at multiples of lcm(level-1 node, d) / d rows: every rank's rows cover whole level-1 nodes of the
    cascade over [n*d], the ranks' node ranges tile [0, n1_nodes), and shapes whose nodes are too coarse fall
    back (None)."""
    import ctypes
    from clane_b200 import _lib
    L = _lib.lib()
    for n, d, world, want in [(169343, 128, 2, True), (169343, 128, 8, True), (2449029, 100, 8, True), (19717, 500, 2, True),
                              (2708, 1433, 2, False), (34, 2, 2, False), (50803, 128, 4, True)]:
        per = cdist.aligned_rows_per_rank(n, d, world)
        assert (per is not None) == want, (n, d, world, per)
        if per is None:
            continue
        nodes, elems = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.clane_cascade_shape(n * d, ctypes.byref(nodes), ctypes.byref(elems)))
        assert (cdist.node_align_rows(n, d) * d) % elems.value == 0 and per % cdist.node_align_rows(n, d) == 0
        assert per % cdist.ROW_ALIGN == 0 and per * world >= n
        prev = 0
        for r in range(world):
            lo, hi = cdist.row_range(n, world, r, per)
            nlo, nhi = cdist.own_node_range(n, d, nodes.value, lo, hi)
            assert nlo == prev and nlo <= nhi
            if hi < n:
                assert (hi * d) % elems.value == 0 and nhi == hi * d // elems.value   # no node straddles the cut
            prev = nhi
        assert prev == nodes.value
        assert max(cdist.row_range(n, world, r, per)[1] - cdist.row_range(n, world, r, per)[0] for r in range(world)) <= 1.1 * n / world + per - n // world


def _worker(rank, world, port, n, ld, nodes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        per = cdist.rows_per_rank(n, world)
        lo, hi = cdist.row_range(n, world, rank)
        rng = np.random.default_rng(0)
        truth = torch.from_numpy(rng.standard_normal((per * world, ld)).astype(np.float32))
        truth[n:] = 0
        # each rank "sweeps" its own rows into an otherwise stale replica, then the slices are exchanged
        mine = torch.full_like(truth, -1.0)
        mine[lo:hi] = truth[lo:hi]
        mine[n:] = 0
        cdist.gather_rows(mine[rank * per:(rank + 1) * per].clone(), mine)
        ok_gather = torch.equal(mine[:n], truth[:n])
        # disjoint cascade slots: each rank fills its node range, all-reduce(SUM) == concatenation, exactly
        slots = torch.from_numpy(rng.standard_normal(((nodes + 2), 32)).astype(np.float32))
        p1 = torch.zeros_like(slots)
        nlo, nhi = cdist.node_range(nodes, world, rank)
        p1[nlo:nhi] = slots[nlo:nhi]
        if rank == world - 1:
            p1[nodes:] = slots[nodes:]          # the trailing partial-node / leftover-row slots
        dist.all_reduce(p1, op=dist.ReduceOp.SUM)
        ok_reduce = torch.equal(p1, slots)
        out.put((rank, ok_gather, ok_reduce))
    finally:
        dist.destroy_process_group()


def test_slice_exchange_and_slot_reduction_gloo():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1003, 12, 37, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
