"""The stage in front of the edit: DDIM inversion of the encoded frames and the on-disk hand-off the drivers read
(reference preprocess.py:198-230 `ddim_inversion`, :232-261 `ddim_sample`, :227-229 / :313-314 the files).

SURVEY.md §8 f-4: this is not part of the hot path — plain UNet forwards, no TokenFlow operator — and the parts of the
reference's preprocess script that need Stable-Diffusion weights (VAE encode / decode, CLIP text encoder, the depth /
ControlNet variants) stay out of scope.  What is here is the latent-space arithmetic and the file format, so that a
latents directory can be produced for, and read back by, `TokenFlowEditor` / the reference drivers
(`tokenflow_utils.load_source_latents_t`): frames are independent in this stage, so with several ranks each rank
inverts its own contiguous share and the saved tensors are all-gathered.
"""
from __future__ import annotations

import os
from typing import Iterable, Optional

import torch


class LatentInverter:
    def __init__(self, unet, scheduler, n_timesteps: int, world_size: int = 1, rank: int = 0, group=None):
        """`scheduler.set_timesteps(n_timesteps)` defines the inversion grid (reference default: 500 steps, of which
        the 50 sampling timesteps are saved)."""
        self.unet, self.scheduler = unet, scheduler
        self.device = next(unet.parameters()).device
        self.scheduler.set_timesteps(n_timesteps, device=self.device)
        self.world_size, self.rank, self.group = world_size, rank, group

    # -- the two DDIM directions -------------------------------------------------------------------------------
    def _alphas(self, t: int, t_prev: Optional[int]):
        a_t = float(self.scheduler.alphas_cumprod[t])
        a_prev = float(self.scheduler.alphas_cumprod[t_prev]) if t_prev is not None else float(self.scheduler.final_alpha_cumprod)
        return a_t ** 0.5, (1 - a_t) ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5          # mu, sigma, mu_prev, sigma_prev

    def _eps(self, x, t: int, cond):
        out = self.unet(x, torch.tensor(t, device=x.device), encoder_hidden_states=cond.repeat(x.shape[0], 1, 1))
        return out["sample"] if isinstance(out, dict) else out.sample

    def _local(self, n: int):
        per = -(-n // self.world_size)
        return self.rank * per, min(n, (self.rank + 1) * per)

    def _gathered(self, x_local, n: int):
        if self.world_size == 1:
            return x_local
        import torch.distributed as dist
        per = -(-n // self.world_size)
        pad = per - x_local.shape[0]
        if pad:
            x_local = torch.cat([x_local, x_local.new_zeros((pad,) + tuple(x_local.shape[1:]))])
        out = torch.empty((self.world_size * per,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
        dist.all_gather_into_tensor(out, x_local.contiguous(), group=self.group)
        return out[:n]

    @torch.no_grad()
    def ddim_inversion(self, cond: torch.Tensor, latent_frames: torch.Tensor, save_path: Optional[str], batch_size: int,
                       save_latents: bool = True, timesteps_to_save: Optional[Iterable[int]] = None) -> torch.Tensor:
        """Reference preprocess.py:198-230.  latent_frames [N,4,h,w] (clean, VAE-encoded) → the latents at the noisiest
        timestep; `noisy_latents_<t>.pt` is written for every t in `timesteps_to_save` (default: all) and for the
        last one."""
        ts = [int(t) for t in reversed(self.scheduler.timesteps.tolist())]                 # ascending noise level
        keep = set(int(t) for t in timesteps_to_save) if timesteps_to_save is not None else set(ts)
        n = latent_frames.shape[0]
        lo, hi = self._local(n)
        x = latent_frames[lo:hi].clone()
        if save_latents and save_path is not None:
            os.makedirs(os.path.join(save_path, "latents"), exist_ok=True)
        for i, t in enumerate(ts):
            mu, sigma, mu_prev, sigma_prev = self._alphas(t, ts[i - 1] if i > 0 else None)
            for b in range(0, x.shape[0], batch_size):
                xb = x[b:b + batch_size]
                eps = self._eps(xb, t, cond)
                pred_x0 = (xb - sigma_prev * eps) / mu_prev
                x[b:b + batch_size] = mu * pred_x0 + sigma * eps
            if save_latents and save_path is not None and (t in keep or i == len(ts) - 1):
                full = self._gathered(x, n)
                if self.rank == 0:
                    torch.save(full, os.path.join(save_path, "latents", f"noisy_latents_{t}.pt"))
        return self._gathered(x, n)

    @torch.no_grad()
    def ddim_sample(self, x: torch.Tensor, cond: torch.Tensor, batch_size: int) -> torch.Tensor:
        """Reference preprocess.py:232-261: deterministic DDIM reconstruction from the inverted latents (the
        `inverted.mp4` check of the reference, in latent space)."""
        ts = [int(t) for t in self.scheduler.timesteps.tolist()]
        n = x.shape[0]
        lo, hi = self._local(n)
        x = x[lo:hi].clone()
        for i, t in enumerate(ts):
            mu, sigma, mu_prev, sigma_prev = self._alphas(t, ts[i + 1] if i < len(ts) - 1 else None)
            for b in range(0, x.shape[0], batch_size):
                xb = x[b:b + batch_size]
                eps = self._eps(xb, t, cond)
                pred_x0 = (xb - sigma * eps) / mu
                x[b:b + batch_size] = mu_prev * pred_x0 + sigma_prev * eps
        return self._gathered(x, n)


def write_inversion_prompt(save_path: str, prompt: str) -> None:
    """Reference preprocess.py:313-314."""
    os.makedirs(save_path, exist_ok=True)
    with open(os.path.join(save_path, "inversion_prompt.txt"), "w") as f:
        f.write(prompt)
