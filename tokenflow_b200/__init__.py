"""tokenflow_b200 — B200-native (sm_100a) implementation of TokenFlow's per-denoise-step hot path.

Layout (only what the path needs):
    csrc/               CUDA kernels + the C-ABI (include/tokenflow_b200.h)
    ops.py              ctypes binding; `CudaOps` = one kernel launch per operator, no fallback
    tokenflow_utils.py  drop-in for the reference hook layer (same names / signatures)
    util.py             drop-in for the names the drivers import from `util`
    sd_unet.py, scheduler.py   diffusers-shaped random-init SD UNet + DDIM (diffusers is not installed)
    editor.py           the caller: mirror of the reference `batched_denoise_step` loop, 1..N GPUs
"""
__version__ = "0.1.0"
