"""CPU tier, build container only: the UNMODIFIED reference driver (`/root/reference/run_tokenflow_pnp.py`
and `run_tokenflow_sdedit.py`, class `TokenFlow`) running on this repo's drop-in `tokenflow_utils` / `util`
modules.  The driver's own `init_method`, `denoise_step` and `batched_denoise_step` are executed as they
are; only `__init__` (Stable-Diffusion download, VAE, CLIP, video files) is bypassed, and `diffusers` —
which is not installed — is a stub that is never called.  The result must equal the golden produced by the
reference hooks (oracle/gen_golden.py), which proves both that the hooks are a drop-in under the reference's
own caller and that `tokenflow_b200/editor.py` mirrors that caller.  The oracle ops stand in for the CUDA
kernels here (no GPU in this tier)."""
import importlib.util
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

from oracle import ref_shim
from oracle.oracle_ops import OracleOps
from tokenflow_b200 import sd_unet
from tokenflow_b200 import tokenflow_utils as tfu
from tokenflow_b200.editor import synthetic_inputs, write_latents_dir
from tokenflow_b200.scheduler import DDIMScheduler

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


def _load_driver(filename):
    """Import a reference driver with `tokenflow_utils` / `util` resolving to this repo's drop-ins."""
    import tokenflow_utils as dropin_tf          # top-level drop-in modules of this repo
    import util as dropin_util
    stub = types.ModuleType("diffusers")
    stub.DDIMScheduler = DDIMScheduler
    stub.StableDiffusionPipeline = type("StableDiffusionPipeline", (), {})
    saved = {k: sys.modules.get(k) for k in ("diffusers", "tokenflow_utils", "util")}
    sys.modules.update({"diffusers": stub, "tokenflow_utils": dropin_tf, "util": dropin_util})
    try:
        name = "_ref_driver_" + filename.replace(".py", "")
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref_shim.REFERENCE_DIR, filename))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)             # `if __name__ == '__main__'` does not fire
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def _make_driver(mod, c, tmp_path, mode):
    cfg = dict(c["config"])
    unet = sd_unet.build_unet("tiny", seed=c["seed"])
    x, text, pnp, src = synthetic_inputs(cfg["n_frames"], c["latent"], unet.config.cross_attention_dim,
                                         cfg["n_timesteps"], seed=c["seed"], ctx_len=c["ctx_len"])
    lat_dir = write_latents_dir(str(tmp_path), src)
    ed = mod.TokenFlow.__new__(mod.TokenFlow)
    nn.Module.__init__(ed)
    ed.config = {"batch_size": cfg["batch_size"], "guidance_scale": cfg["guidance_scale"], "n_frames": cfg["n_frames"],
                 "n_timesteps": cfg["n_timesteps"]}
    ed.device = "cpu"
    ed.sd_version = "1.5"
    ed.unet = unet
    ed.scheduler = DDIMScheduler()
    ed.scheduler.set_timesteps(cfg["n_timesteps"], device="cpu")
    if mode == "sdedit":                          # run_tokenflow_sdedit.py:57
        ed.scheduler.timesteps = ed.scheduler.timesteps[int(1 - cfg["start"] * cfg["n_timesteps"]):]
    ed.latents_path = lat_dir
    ed.text_embeds = text
    ed.pnp_guidance_embeds = pnp
    return ed, x, cfg


@pytest.mark.parametrize("mode,driver,golden", [("pnp", "run_tokenflow_pnp.py", "unet_c1_pnp.pt"),
                                                ("sdedit", "run_tokenflow_sdedit.py", "unet_c1_sdedit.pt")])
def test_unmodified_reference_driver_on_dropin_hooks(mode, driver, golden, golden_dir, tmp_path):
    c = torch.load(os.path.join(golden_dir, golden), weights_only=False)
    mod = _load_driver(driver)
    assert mod.register_pivotal is tfu.register_pivotal          # `from tokenflow_utils import *` bound OUR hooks
    tfu._install_ops_for_testing(OracleOps())
    ed, x, cfg = _make_driver(mod, c, tmp_path, mode)
    if mode == "pnp":                                            # run_tokenflow_pnp.py:253-256
        ed.init_method(conv_injection_t=int(cfg["n_timesteps"] * cfg["pnp_f_t"]),
                       qk_injection_t=int(cfg["n_timesteps"] * cfg["pnp_attn_t"]))
    else:                                                        # run_tokenflow_sdedit.py:191-193
        ed.init_method()
    assert [int(t) for t in ed.scheduler.timesteps] == c["timesteps"]
    torch.manual_seed(c["seed"])
    indices = torch.arange(cfg["n_frames"])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                          # the driver's cuda autocast decorator on a CPU box
        for i, t in enumerate(ed.scheduler.timesteps):           # the body of sample_loop (:266-267), VAE decode omitted
            x = ed.batched_denoise_step(x, t, indices)
            assert torch.allclose(x, c["steps"][i], atol=2e-4, rtol=1e-4), f"step {i}"
    assert torch.allclose(x, c["out"], atol=2e-4, rtol=1e-4)
