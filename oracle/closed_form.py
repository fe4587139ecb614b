"""ORACLE (test infrastructure) — closed-form numpy/fp64 statement of the hot path (SURVEY.md
Appendix A).  Independent of torch; pure loops over frames/heads, used only at small sizes to
cross-check `oracle/tokenflow_oracle.py` and the golden vectors.

Reference lines restated: tokenflow_utils.py:114-199 / :224-281 (extended attention),
:329-348 + util.py:61-69 (NN field), :361-393 (propagation).
"""
from __future__ import annotations

import numpy as np


def _softmax(x: np.ndarray) -> np.ndarray:
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def extended_attention(q, k, v, heads: int, scale: float, inject: bool = False) -> np.ndarray:
    """q,k,v [3n,S,dim] → [3n,S,dim] (before to_out).  src frame f: own keys; uncond/cond frame f:
    keys of all n frames of that stream, frame-major."""
    q, k, v = (np.asarray(t, dtype=np.float64) for t in (q, k, v))
    b, S, dim = q.shape
    n, d = b // 3, dim // heads
    if inject:  # :124-130
        q = q.copy(); k = k.copy()
        q[n:2 * n] = q[:n]; q[2 * n:] = q[:n]
        k[n:2 * n] = k[:n]; k[2 * n:] = k[:n]
    out = np.zeros_like(q)
    for s in range(3):
        for f in range(n):
            for j in range(heads):
                c = slice(j * d, (j + 1) * d)
                Q = q[s * n + f][:, c]
                if s == 0:
                    Kk, Vv = k[f][:, c], v[f][:, c]
                else:
                    Kk = k[s * n:(s + 1) * n][:, :, c].reshape(n * S, d)
                    Vv = v[s * n:(s + 1) * n][:, :, c].reshape(n * S, d)
                out[s * n + f][:, c] = _softmax(Q @ Kk.T * scale) @ Vv
    return out


def nn_index(x, y) -> np.ndarray:
    """argmax_c cos(x_r, y_c), first index on ties.  x [R,dim], y [C,dim] → int64 [R]."""
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    x = x / np.linalg.norm(x, axis=-1, keepdims=True)
    y = y / np.linalg.norm(y, axis=-1, keepdims=True)
    return np.argmax(x @ y.T, axis=-1)


def blend_weight(f: int, B: int) -> float:
    """σ(d2/(d1+d2)) with d1=|g-(iB+B//2)|, d2=|g-((i-1)B+B//2)|, g=iB+f → independent of i."""
    d1 = abs(f - B // 2)
    d2 = abs(f + B - B // 2)
    return 1.0 / (1.0 + np.exp(-(d2 / (d1 + d2))))


def propagate(A, idx1, idx2, batch_idx: int, B: int) -> np.ndarray:
    """A [3,K,S,dim]; idx [B*S] → [3B,S,dim]."""
    A = np.asarray(A, dtype=np.float64)
    _, K, S, dim = A.shape
    out = np.zeros((3, B, S, dim))
    idx1 = np.asarray(idx1).reshape(B, S)
    idx2 = None if idx2 is None else np.asarray(idx2).reshape(B, S)
    for s in range(3):
        for f in range(B):
            if idx2 is None:
                out[s, f] = A[s, batch_idx][idx1[f]]
            else:
                w = blend_weight(f, B)
                out[s, f] = w * A[s, batch_idx][idx1[f]] + (1 - w) * A[s, batch_idx - 1][idx2[f]]
    return out.reshape(3 * B, S, dim)
