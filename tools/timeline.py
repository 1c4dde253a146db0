"""Timeline of the kernels of one batch of sweeps (clane_plan_trace): when each kernel's first CTA started and its last
CTA ended, relative to the first stamp.  python tools/timeline.py [workload] [sweeps]"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clane_b200 import _lib, similarity, synth  # noqa: E402
from clane_b200.graph import Graph  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n, src, dst, X = synth.make_graph(name, seed=0)
g = Graph.from_arrays(n, src, dst, X)
S = g._device_state()
g._build_P_device(similarity.CosineSimilarity())
L = _lib.lib()
sh = S.stream.cuda_stream
gamma = ctypes.c_float(float(np.float32(0.76)))


def run(cnt):
    _lib.check(L.clane_patience_reset(S.state.data_ptr(), 1 << 30, 0, sh))
    _lib.check(L.clane_sweeps(S.plan.handle, S.X.data_ptr(), S.Zptrs, S.cur, S.rowptr.data_ptr(), S.col.data_ptr(),
                              S.w.data_ptr(), gamma, cnt, 0, S.state.data_ptr(), S.log.data_ptr(), S.log_cap, sh))
    S.cur = (S.cur + cnt) % 3
    torch.cuda.synchronize()


run(12)
_lib.check(L.clane_plan_trace(S.plan.handle, 1, None))
run(k)          # captured (first replay) ...
_lib.check(L.clane_plan_trace(S.plan.handle, 1, None))
run(k)          # ... and replayed: this is the one we read
buf = np.zeros(64 * 6 * 2, np.uint64)
_lib.check(L.clane_plan_trace(S.plan.handle, 0, buf.ctypes.data))
T = buf.reshape(64, 6, 2)[:k].astype(np.int64)
names = ["segments", "chain_long", "chain_short", "spans", "l1_tail", "finish"]
valid = T[:, :, 1] > 0
t0 = T[:, :, 0][valid].min()
for t in range(k):
    row = [f"sweep {t}:"]
    for s_, nm in enumerate(names):
        if T[t, s_, 1] > 0:
            row.append(f"{nm} {(T[t, s_, 0] - t0) / 1e3:8.1f}-{(T[t, s_, 1] - t0) / 1e3:8.1f}")
    print("  ".join(row))
print(f"per sweep: {(T[:, :, 1].max() - t0) / 1e3 / k:.1f} us")
