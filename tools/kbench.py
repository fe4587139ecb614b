"""Kernel micro-benchmarks at the BASELINE C2 shapes (CUDA events, L2 flushed between timed
iterations).  Usage: python tools/kbench.py [--json out.json]"""
import argparse
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokenflow_b200.ops import CudaOps, blend_weights  # noqa: E402


def timed(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                 # write > L2 capacity ...
            flush[:: 64].sum()            # ... then touch it read-only so no dirty lines are evicted mid-measurement
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    ops = CudaOps()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    res = {}
    levels = [(4096, 320, 8), (1024, 640, 8), (256, 1280, 8), (64, 1280, 8)]
    n, B, K = 5, 8, 5
    for S, dim, heads in levels:
        d = dim // heads
        torch.manual_seed(0)
        q, k, v = (torch.randn(3 * n, S, dim, device="cuda").half() for _ in range(3))
        for inject in (False,):
            ms, best = timed(lambda: ops.ext_attn(q, k, v, heads, d ** -0.5, inject), flush=flush)
            flops = 4 * K * S * S * dim * (2 * K + 1)
            res[f"ext_attn_S{S}_d{d}"] = {"ms": ms, "best_ms": best, "tflops": flops / ms / 1e9}
        # SDPA comparison (cuDNN/flash backend): same contraction, K/V materialised
        qh = q.view(3 * n, S, heads, d).transpose(1, 2)
        kh = k.view(3 * n, S, heads, d).transpose(1, 2)
        vh = v.view(3 * n, S, heads, d).transpose(1, 2)

        def sdpa():
            o0 = torch.nn.functional.scaled_dot_product_attention(qh[:n], kh[:n], vh[:n])
            outs = [o0]
            for s in (1, 2):
                kk = kh[s * n:(s + 1) * n].permute(1, 0, 2, 3).reshape(1, heads, n * S, d).expand(n, -1, -1, -1)
                vv = vh[s * n:(s + 1) * n].permute(1, 0, 2, 3).reshape(1, heads, n * S, d).expand(n, -1, -1, -1)
                outs.append(torch.nn.functional.scaled_dot_product_attention(qh[s * n:(s + 1) * n], kk, vv))
            return outs
        try:
            ms, best = timed(sdpa, flush=flush)
            res[f"sdpa_S{S}_d{d}"] = {"ms": ms, "best_ms": best, "tflops": 4 * K * S * S * dim * (2 * K + 1) / ms / 1e9}
        except Exception as ex:  # noqa: BLE001
            res[f"sdpa_S{S}_d{d}"] = {"error": str(ex)[:200]}

        x = torch.randn(B, S, dim, device="cuda")
        piv = torch.randn(K, S, dim, device="cuda")
        ms, best = timed(lambda: ops.unit_rows(x), flush=flush)
        res[f"unit_rows_S{S}_dim{dim}"] = {"ms": ms, "best_ms": best, "gbs": B * S * dim * 6 / ms / 1e6}
        xu, pu = ops.unit_rows(x), ops.unit_rows(piv)
        kf_a, kf_b = [2] * B, [1] * B
        ms, best = timed(lambda: ops.nn_field(xu, pu, kf_a, kf_b), flush=flush)
        res[f"nn_field_S{S}_dim{dim}"] = {"ms": ms, "best_ms": best, "tflops": 2 * B * S * 2 * S * dim / ms / 1e9}

        def cublas_nn():
            sim = xu.view(-1, dim) @ pu[[2, 1]].reshape(-1, dim).T
            s1, s2 = sim.chunk(2, dim=1)
            return s1.argmax(-1), s2.argmax(-1)
        ms, best = timed(cublas_nn, flush=flush)
        res[f"cublas_argmax_S{S}_dim{dim}"] = {"ms": ms, "best_ms": best, "tflops": 2 * B * S * 2 * S * dim / ms / 1e9}

        A = torch.randn(3, K, S, dim, device="cuda").half()
        idx_a, idx_b = ops.nn_field(xu, pu, kf_a, kf_b)
        resid = torch.randn(3 * B, S, dim, device="cuda").half()
        w = blend_weights(B)
        ms, best = timed(lambda: ops.propagate(A, idx_a, idx_b, kf_a, kf_b, w, resid), flush=flush)
        # algorithmic bytes: out write + residual read + 2 keyframe slabs x 3 streams + indices
        byts = 3 * B * S * dim * 2 * 2 + 3 * 2 * S * dim * 2 + 2 * 4 * B * S
        res[f"propagate_S{S}_dim{dim}"] = {"ms": ms, "best_ms": best, "gbs": byts / ms / 1e6, "bytes": byts}
    for k_, v_ in res.items():
        print(k_, json.dumps(v_))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
