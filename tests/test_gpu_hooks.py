"""GPU tier: the drop-in hook layer end to end on the CUDA kernels.

The plumbing of the hooks is proven equal to the unmodified reference on CPU
(tests/test_hooks_cpu.py).  Here the same hooks run under `torch.autocast(float16)` on the GPU twice:
once on the CUDA kernels (product) and once with the oracle ops installed — which, under autocast
on the same device, launches exactly the library kernels the reference's GPU path launches — and
the two are compared."""
import os

import pytest
import torch

from oracle.oracle_ops import OracleOps
from tokenflow_b200 import sd_unet
from tokenflow_b200 import tokenflow_utils as tfu
from tokenflow_b200.editor import TokenFlowEditor, synthetic_inputs
from tokenflow_b200.scheduler import DDIMScheduler

pytestmark = pytest.mark.gpu


def _run(ops, mode, n_frames=4, batch=2, steps=2, latent=16, seed=1, kind="tiny", fused=False):
    tfu._install_ops_for_testing(ops)
    unet = sd_unet.build_unet(kind, seed=seed, device="cuda", dtype=torch.float16)
    cfg = {"n_frames": n_frames, "batch_size": batch, "n_timesteps": steps, "guidance_scale": 7.5,
           "mode": mode, "pnp_attn_t": 0.5, "pnp_f_t": 0.8, "start": 0.9, "fused_pass": fused}
    x, text, pnp, src = synthetic_inputs(n_frames, latent, unet.config.cross_attention_dim, steps, seed=seed,
                                         device="cuda", dtype=torch.float16, ctx_len=7)
    ed = TokenFlowEditor(unet, DDIMScheduler(), tfu, cfg, text, pnp, source_latents=lambda t: src[t])
    ed.init_method()
    torch.manual_seed(seed)
    idx_log = []
    out = ed.sample_loop(x)
    blocks = tfu._transformer_blocks(ed)
    for b in blocks:
        idx_log.append(tuple(None if t is None else t.clone().long().reshape(-1) for t in b._tf_nn_idx))
    return out.float(), ed.keyframe_log, idx_log


@pytest.mark.parametrize("mode,steps", [("pnp", 2), ("sdedit", 10)])
def test_tiny_unet_edit_cuda_vs_reference_gpu_path(mode, steps):
    want, kf_w, idx_w = _run(OracleOps(), mode, steps=steps)
    got, kf_g, idx_g = _run(None, mode, steps=steps)             # None -> product CudaOps
    assert kf_g == kf_w
    assert torch.isfinite(got).all()
    rel = (got - want).norm() / want.norm()
    assert rel.item() < 2e-2, rel.item()


def test_golden_c1_through_cuda(golden_dir):
    """BASELINE config C1 (miniature) golden from the unmodified reference (fp32 CPU) vs the CUDA
    product path (fp16 autocast): same keyframes, outputs equal to fp16 accumulation error."""
    c = torch.load(os.path.join(golden_dir, "unet_c1_pnp.pt"), weights_only=False)
    got, kf, _ = _run(None, "pnp", steps=c["config"]["n_timesteps"], seed=c["seed"], latent=c["latent"])
    assert kf == c["keyframes"]
    rel = (got.cpu() - c["out"]).norm() / c["out"].norm()
    assert rel.item() < 5e-2, rel.item()


def test_strict_dtype_env(monkeypatch):
    """TOKENFLOW_B200_STRICT_DTYPE=1 emits the reference's promoted fp32 hidden state."""
    from tokenflow_b200.ops import CudaOps, blend_weights
    ops = CudaOps()
    A = torch.randn(3, 2, 16, 8, device="cuda").half()
    idx = torch.randint(0, 16, (2, 16), device="cuda", dtype=torch.int32)
    out = ops.propagate(A, idx, idx, [1, 1], [0, 0], blend_weights(2), None, out_dtype=torch.float32)
    assert out.dtype == torch.float32


def test_sd21_shape_block_cuda_vs_reference_gpu_path():
    """BASELINE C4 shapes at the hook level: one SD2.1-shape transformer block (dim 320, 5 heads x 64) at
    2304 tokens, SDEdit flavour (no injection): pivotal pass over 3 keyframes + two frame passes, CUDA ops
    vs the oracle ops under the same autocast."""
    import torch.nn as nn

    class _U(nn.Module):
        def __init__(self, block):
            super().__init__()
            site = nn.Module()
            site.transformer_blocks = nn.ModuleList([block])
            ups = []
            for _ in range(4):
                u = nn.Module()
                u.attentions = nn.ModuleList([site, site, site])
                ups.append(u)
            self.up_blocks = nn.ModuleList(ups)

    class _W(nn.Module):
        def __init__(self, unet):
            super().__init__()
            self.unet = unet

    def run(ops):
        tfu._install_ops_for_testing(ops)
        torch.manual_seed(3)
        block = sd_unet.BasicTransformerBlock(320, 5, 64, 1024).cuda().half().eval()
        model = _W(_U(block))
        tfu.register_extended_attention(model)
        tfu.set_tokenflow(model.unet)
        K, B, S = 3, 2, 2304
        outs = []
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            h = torch.randn(3 * K, S, 320, device="cuda").half()
            ctx = torch.randn(3 * K, 77, 1024, device="cuda").half()
            tfu.register_pivotal(model, True)
            outs.append(block(h, encoder_hidden_states=ctx).float())
            tfu.register_pivotal(model, False)
            for i in (0, 2):
                hf = (h[:K][i].unsqueeze(0).repeat(B, 1, 1) + 0.3 * torch.randn(B, S, 320, device="cuda").half())
                hf = torch.cat([hf, torch.randn(2 * B, S, 320, device="cuda").half()])
                tfu.register_batch_idx(model, i)
                outs.append(block(hf, encoder_hidden_states=ctx[:3 * B]).float())
                outs.append(block._tf_nn_idx[0].long().reshape(-1).clone())
        return outs

    want = run(OracleOps())
    got = run(None)
    assert torch.allclose(got[0], want[0], atol=3e-3, rtol=3e-3)          # pivotal pass
    for j in (1, 3):                                                       # frame passes: (output, NN indices)
        g_out, w_out, g_idx, w_idx = got[j], want[j], got[j + 1], want[j + 1]
        same = (g_idx == w_idx)
        assert (~same).float().mean().item() < 5e-3                       # fp16 tie classes only
        rows = same.view(1, -1).expand(3, -1).reshape(-1)                  # [3*B*S] rows whose NN index agrees
        g2, w2 = g_out.reshape(-1, 320)[rows], w_out.reshape(-1, 320)[rows]
        assert torch.allclose(g2, w2, atol=3e-3, rtol=3e-3)


@pytest.mark.parametrize("mode,steps", [("pnp", 2), ("sdedit", 10)])
def test_fused_pass_cuda(mode, steps):
    """One UNet call per step ([pivotal samples | frames]) on the CUDA kernels == separate passes."""
    want, kf_w, _ = _run(None, mode, steps=steps)
    got, kf_g, _ = _run(None, mode, steps=steps, fused=True)
    assert kf_g == kf_w and torch.isfinite(got).all()
    rel = (got - want).norm() / want.norm()
    assert rel.item() < 2e-2, rel.item()
