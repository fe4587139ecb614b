"""Launch each hot-path kernel a few times at the BASELINE C2 top-level shapes (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenflow_b200.ops import CudaOps, blend_weights  # noqa: E402

ops = CudaOps()
n, B, K, S, dim, heads = 5, 8, 5, 4096, 320, 8
d = dim // heads
torch.manual_seed(0)
q, k, v = (torch.randn(3 * n, S, dim, device="cuda").half() for _ in range(3))
x = torch.randn(B, S, dim, device="cuda")
piv = torch.randn(K, S, dim, device="cuda")
A = torch.randn(3, K, S, dim, device="cuda").half()
resid = torch.randn(3 * B, S, dim, device="cuda").half()
kf_a, kf_b, w = [2] * B, [1] * B, blend_weights(B)
norm = torch.nn.LayerNorm(dim).cuda().half()
hid = torch.randn(B, S, dim, device="cuda").half()
S2, dim2 = 1024, 640                                     # SD1.5 middle level (d = 80): the two-half kernel
q2, k2, v2 = (torch.randn(3 * n, S2, dim2, device="cuda").half() for _ in range(3))
hid_piv = torch.randn(3 * n, S, dim, device="cuda").half()
for _ in range(3):
    ops.ext_attn(q, k, v, heads, d ** -0.5, False)       # quad-stream kernel, all 15 samples
    ops.ext_attn(q, k, v, heads, d ** -0.5, True)        # PnP injection: paired kernel (uncond + cond) + source samples
    ops.ext_attn(q2, k2, v2, heads, (dim2 // heads) ** -0.5, False)
    ops.layernorm_rows(hid_piv, norm, n)
    ops.layernorm_unit_rows(hid, norm)
    xu, pu = ops.unit_rows(x), ops.unit_rows(piv)
    idx_a, idx_b = ops.nn_field(xu, pu, kf_a, kf_b)
    ops.propagate(A, idx_a, idx_b, kf_a, kf_b, w, resid)
torch.cuda.synchronize()
print("done")
