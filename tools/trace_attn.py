"""Dump the clock64 event trace of CTA 0 of the ping-pong attention kernel (needs a -DTF_TRACE build:
TF_BUILD_TRACE=1 python -m tokenflow_b200._build --force)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenflow_b200.ops import CudaOps  # noqa: E402

ops = CudaOps()
n, S, heads, d = 5, 4096, 8, 40
dim = heads * d
torch.manual_seed(0)
q, k, v = (torch.randn(3 * n, S, dim, device="cuda").half() for _ in range(3))
for _ in range(2):
    ops.ext_attn(q, k, v, heads, d ** -0.5, False)
torch.cuda.synchronize()
T, E = 40, 8
buf = (ctypes.c_longlong * (3 * T * E))()
ops.lib.tf_debug_read_attn_trace.restype = ctypes.c_int
assert ops.lib.tf_debug_read_attn_trace(buf, 3 * T * E) == 0
arr = [[[buf[(r * T + t) * E + e] for e in range(E)] for t in range(T)] for r in range(3)]
t0 = arr[0][8][0]
print("softmax A/B events: 0 wait_s  1 s_ready  2 ld_done  3 max_done  4 exp+st_issued  5 st_done  6 arrived   (cycles, relative)")
for t in range(8, 20):
    for r in (0, 1):
        ev = arr[r][t]
        print(f"t={t:2d} {'AB'[r]}: " + " ".join(f"{e - t0:7d}" for e in ev[:7]) + "   | d: " + " ".join(f"{ev[i + 1] - ev[i]:5d}" for i in range(6)))
print("MMA events per tile: A[wait_p seen_p pv_issued qk_issued] B[...]")
for t in range(8, 20):
    ev = arr[2][t]
    print(f"t={t:2d} M: " + " ".join(f"{e - t0:7d}" for e in ev[:8]))
