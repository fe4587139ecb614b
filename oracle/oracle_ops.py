"""ORACLE (test infrastructure) — an op object with the interface of `tokenflow_b200.ops.CudaOps`
whose arithmetic is the oracle restatement (`oracle/tokenflow_oracle.py`).

Installed through `tokenflow_utils._install_ops_for_testing` by `tests/` (hook plumbing on CPU,
BASELINE config C1) and by `bench.py --impl reference` / the `cpu_baseline` leg (the reference's
algorithm timed on host cores).  Never imported by the product package.

`unit_rows` deliberately returns the *un-normalised* rows: the reference normalises inside
`batch_cosine_sim` (util.py:66-67) in whatever dtype it is running (fp32 on CPU), so the oracle's
`nn_field` does the same instead of rounding to fp16 first.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import tokenflow_oracle as O


class OracleOps:
    name = "oracle-cpu"

    def launch_count(self) -> int:
        return 0

    def unit_rows(self, x: torch.Tensor) -> torch.Tensor:
        return x

    def layernorm_unit_rows(self, x: torch.Tensor, norm) -> torch.Tensor:
        return norm(x)          # reference :323; the L2 normalisation happens inside nn_field (util.py:66-67)

    def nn_field(self, x: torch.Tensor, piv: torch.Tensor, kf_a: Sequence[int], kf_b: Sequence[int]
                 ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        F_, S, dim = x.shape
        idx_a = torch.empty((F_, S), dtype=torch.int64, device=x.device)
        any_b = any(int(b) >= 0 for b in kf_b)
        idx_b = torch.zeros((F_, S), dtype=torch.int64, device=x.device) if any_b else None
        # group frames that share (kf_a, kf_b): the reference's per-batch call is one such group
        groups = {}
        for f, key in enumerate(zip(kf_a, kf_b)):
            groups.setdefault((int(key[0]), int(key[1])), []).append(f)
        for (a, b), frames in groups.items():
            xs = x[frames].reshape(-1, dim)
            idx_a[frames] = O.cosine_sim(xs, piv[a]).argmax(dim=-1).view(len(frames), S)
            if b >= 0:
                idx_b[frames] = O.cosine_sim(xs, piv[b]).argmax(dim=-1).view(len(frames), S)
        return idx_a, idx_b

    def propagate(self, A: torch.Tensor, idx_a: torch.Tensor, idx_b: Optional[torch.Tensor],
                  kf_a: Sequence[int], kf_b: Sequence[int], w: Sequence[float],
                  residual: Optional[torch.Tensor], out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        _, K, S, dim = A.shape
        F_ = idx_a.shape[0]
        outs = []
        for f in range(F_):
            a1 = A[:, int(kf_a[f])][:, idx_a[f].long()]                 # [3, S, dim]
            if int(kf_b[f]) >= 0:
                a2 = A[:, int(kf_b[f])][:, idx_b[f].long()]
                # a dimensioned fp32 weight tensor, like the reference's w1.repeat(3,1,S,dim) (:385):
                # it promotes the fp16 rows to fp32 (a 0-dim tensor would not)
                wf = torch.full((1, 1, 1), float(w[f]), dtype=torch.float32, device=A.device)
                a1 = wf * a1 + (1 - wf) * a2                            # reference :388
            outs.append(a1)
        out = torch.stack(outs, dim=1).reshape(3 * F_, S, dim)
        if residual is not None:
            out = out + residual.reshape(3 * F_, S, dim)                # reference :397
        if out_dtype is not None:
            out = out.to(out_dtype)
        return out

    def ext_attn(self, q, k, v, heads: int, scale: float, inject: bool) -> torch.Tensor:
        return O.extended_attention(q, k, v, heads, scale, inject)

    def ext_attn_table(self, q, k, v, table, heads: int, scale: float, row0: int = 0, nrows=None) -> torch.Tensor:
        """Sharded-pass form: output sample j attends with q[qs] to the nkv consecutive slabs
        k[k0:k0+nkv], v[v0:v0+nkv] (frame-major), per head — the same per-head softmax(q k^T scale) v
        as reference :173-179."""
        _, S, dim = q.shape
        d = dim // heads
        nrows = S if nrows is None else int(nrows)
        r1 = min(S, row0 + nrows)                      # query tokens [row0, r1); rows past S stay zero
        outs = []
        for (qs, k0, v0, nkv) in table:
            qq = q[qs, row0:r1].reshape(max(0, r1 - row0), heads, d).permute(1, 0, 2)
            kk = k[k0:k0 + nkv].reshape(nkv * S, heads, d).permute(1, 0, 2)
            vv = v[v0:v0 + nkv].reshape(nkv * S, heads, d).permute(1, 0, 2)
            sim = torch.bmm(qq, kk.transpose(-1, -2)) * scale
            o = torch.bmm(sim.softmax(dim=-1), vv)
            o = o.permute(1, 0, 2).reshape(max(0, r1 - row0), dim)
            if o.shape[0] < nrows:
                o = torch.cat([o, o.new_zeros(nrows - o.shape[0], dim)])
            outs.append(o)
        return torch.stack(outs)
