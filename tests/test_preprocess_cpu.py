"""CPU tier: the latent-space part of the reference's preprocess stage (DDIM inversion, reconstruction, file format)."""
import os

import torch

from tokenflow_b200 import sd_unet, tokenflow_utils as tfu
from tokenflow_b200.preprocess import LatentInverter, write_inversion_prompt
from tokenflow_b200.scheduler import DDIMScheduler


class _LinearEps(torch.nn.Module):
    """eps(x, t) = k * x: closed forms exist for both DDIM directions."""

    def __init__(self, k):
        super().__init__()
        self.k = torch.nn.Parameter(torch.tensor(float(k)))

    def forward(self, x, t, encoder_hidden_states=None):
        return {"sample": self.k * x}


def test_inversion_and_sampling_match_the_closed_form_for_linear_eps(tmp_path):
    k, steps = 0.3, 20
    inv = LatentInverter(_LinearEps(k), DDIMScheduler(), steps)
    sch = inv.scheduler
    x0 = torch.randn(5, 4, 8, 8)
    cond = torch.zeros(1, 7, 16)
    ts_up = [int(t) for t in reversed(sch.timesteps.tolist())]
    save = [ts_up[3], ts_up[10]]
    xT = inv.ddim_inversion(cond, x0, str(tmp_path), batch_size=2, timesteps_to_save=save)
    # closed form: every step multiplies the latents by mu * (1 - sigma_prev k) / mu_prev + sigma k  (preprocess.py:207-224)
    a = sch.alphas_cumprod
    factor, at = 1.0, {}
    for i, t in enumerate(ts_up):
        a_t = float(a[t]); a_p = float(a[ts_up[i - 1]]) if i > 0 else float(sch.final_alpha_cumprod)
        factor *= (a_t ** 0.5) * (1 - (1 - a_p) ** 0.5 * k) / a_p ** 0.5 + (1 - a_t) ** 0.5 * k
        at[t] = factor
    assert torch.allclose(xT, x0 * factor, rtol=1e-4, atol=1e-5)
    # files: the requested timesteps + the last one, in the format load_source_latents_t reads (tokenflow_utils.py:43-47)
    files = sorted(os.listdir(tmp_path / "latents"))
    assert files == sorted(f"noisy_latents_{t}.pt" for t in set(save + [ts_up[-1]]))
    for t in save:
        got = tfu.load_source_latents_t(t, str(tmp_path / "latents"))
        assert got.shape == x0.shape and torch.allclose(got, x0 * at[t], rtol=1e-4, atol=1e-5)
    # the reverse direction undoes it step by step for this eps model only approximately (DDIM inversion is not exact);
    # its own closed form: product of mu_prev * (1 - sigma k) / mu + sigma_prev k  (preprocess.py:244-260)
    rec = inv.ddim_sample(xT, cond, batch_size=3)
    ts_dn = [int(t) for t in sch.timesteps.tolist()]
    g = 1.0
    for i, t in enumerate(ts_dn):
        a_t = float(a[t]); a_p = float(a[ts_dn[i + 1]]) if i < len(ts_dn) - 1 else float(sch.final_alpha_cumprod)
        g *= (a_p ** 0.5) * (1 - (1 - a_t) ** 0.5 * k) / a_t ** 0.5 + (1 - a_p) ** 0.5 * k
    assert torch.allclose(rec, xT * g, rtol=1e-4, atol=1e-5)
    write_inversion_prompt(str(tmp_path), "a prompt")
    assert (tmp_path / "inversion_prompt.txt").read_text() == "a prompt"


def test_inversion_feeds_the_editor_on_the_toy_unet(tmp_path):
    """Inverted latents written by LatentInverter are what TokenFlowEditor / the drivers load as source latents."""
    from oracle.oracle_ops import OracleOps
    from tokenflow_b200.editor import TokenFlowEditor, synthetic_inputs
    steps = 4
    unet = sd_unet.build_unet("tiny", seed=1)
    x, text, pnp, _ = synthetic_inputs(4, 16, unet.config.cross_attention_dim, steps, seed=1, ctx_len=7)
    inv = LatentInverter(unet, DDIMScheduler(), steps)
    noisy = inv.ddim_inversion(pnp, x, str(tmp_path), batch_size=2)
    assert torch.isfinite(noisy).all()
    tfu._install_ops_for_testing(OracleOps())
    cfg = {"n_frames": 4, "batch_size": 2, "n_timesteps": steps, "guidance_scale": 7.5, "mode": "pnp",
           "latents_path": str(tmp_path / "latents")}
    ed = TokenFlowEditor(unet, DDIMScheduler(), tfu, cfg, text, pnp)
    ed.init_method()
    torch.manual_seed(1)
    out = ed.sample_loop(noisy.clone())
    assert out.shape == x.shape and torch.isfinite(out).all()
    assert sorted(os.listdir(tmp_path / "latents")) == sorted(f"noisy_latents_{t}.pt" for t in ed._t_host)
