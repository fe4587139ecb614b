"""The caller of the hot path: the denoising loop of the reference drivers
(run_tokenflow_pnp.py:195-240, 264-273; run_tokenflow_sdedit.py:154-205), without the parts that
need Stable-Diffusion weights (VAE, CLIP, image/video I/O — out of scope).

`TokenFlowEditor` is parameterised by the hook module, so the same loop runs
  * this package's hooks (`tokenflow_b200.tokenflow_utils`, CUDA kernels) — the product,
  * the unmodified reference hooks (via oracle/ref_shim.py) — golden-vector generation,
  * this package's hooks with oracle ops installed — CPU plumbing tests / CPU baseline.

Per denoising step (reference batched_denoise_step, :220-233):
  1. draw one random keyframe per batch of B frames (CPU generator, like the reference);
  2. pivotal pass: UNet over [source | uncond | cond] x K keyframes, output discarded — it only
     fills the per-block caches (pivot features, extended-attention outputs);
  3. frame passes: UNet over each batch of B frames; self-attention is replaced by NN propagation;
  4. classifier-free guidance + DDIM update per batch.
"""
from __future__ import annotations

import contextlib
import os
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn


class TokenFlowEditor(nn.Module):
    def __init__(self, unet: nn.Module, scheduler, hooks, config: Dict, text_embeds: torch.Tensor,
                 pnp_guidance_embeds: torch.Tensor, source_latents: Optional[Callable[[int], torch.Tensor]] = None,
                 world_size: int = 1, rank: int = 0, group=None):
        """config keys (names follow configs/config_pnp.yaml): n_frames, batch_size, n_timesteps,
        guidance_scale, mode ('pnp' | 'sdedit'), pnp_attn_t, pnp_f_t, start (sdedit), latents_path.
        text_embeds: [2, L, C] (uncond, cond);  pnp_guidance_embeds: [1, L, C] (inversion prompt)."""
        super().__init__()
        self.unet = unet
        self.scheduler = scheduler
        self.hooks = hooks
        self.config = dict(config)
        self.text_embeds = text_embeds
        self.pnp_guidance_embeds = pnp_guidance_embeds
        self.latents_path = self.config.get("latents_path")
        self._source_latents = source_latents
        self.device = next(unet.parameters()).device
        self.scheduler.set_timesteps(self.config["n_timesteps"], device=self.device)
        if self.config.get("mode", "pnp") == "sdedit":        # run_tokenflow_sdedit.py:57
            start = float(self.config.get("start", 0.9))
            self.scheduler.timesteps = self.scheduler.timesteps[int(1 - start * self.config["n_timesteps"]):]
        self.keyframe_log = []
        self.world_size, self.rank, self.group = world_size, rank, group
        self._src_override = None
        # host ints and device scalars of the timesteps, resolved once: the per-step code never has to read
        # a device tensor back (`int(t)` on a CUDA tensor is a full stream synchronisation)
        self._t_host = [int(t) for t in self.scheduler.timesteps]
        self._t_dev = {t: torch.tensor(t, device=self.device) for t in self._t_host}

    # ------------------------------------------------------------------------------------
    def init_method(self):
        """run_tokenflow_pnp.py:235-240 / run_tokenflow_sdedit.py:191-193."""
        h = self.hooks
        if self.config.get("mode", "pnp") == "pnp":
            n = self.config["n_timesteps"]
            qk_t = int(n * self.config.get("pnp_attn_t", 0.5))
            conv_t = int(n * self.config.get("pnp_f_t", 0.8))
            self.qk_injection_timesteps = self.scheduler.timesteps[:qk_t] if qk_t >= 0 else []
            self.conv_injection_timesteps = self.scheduler.timesteps[:conv_t] if conv_t >= 0 else []
            h.register_extended_attention_pnp(self, self.qk_injection_timesteps)
            h.register_conv_injection(self, self.conv_injection_timesteps)
        else:
            h.register_extended_attention(self)
        h.set_tokenflow(self.unet)

    def source_latents_t(self, t: int) -> torch.Tensor:
        if self._src_override is not None and self._src_override[0] == int(t):
            return self._src_override[1]
        if self._source_latents is not None:
            return self._source_latents(t)
        return self.hooks.load_source_latents_t(t, self.latents_path)

    # ------------------------------------------------------------------------------------
    @torch.no_grad()
    def denoise_step(self, x, t, indices):
        """run_tokenflow_pnp.py:195-218."""
        source_latents = self.source_latents_t(int(t))[indices].to(x.device, x.dtype)
        latent_model_input = torch.cat([source_latents] + ([x] * 2))
        self.hooks.register_time(self, int(t))
        text_embed_input = torch.cat([self.pnp_guidance_embeds.repeat(len(indices), 1, 1),
                                      torch.repeat_interleave(self.text_embeds, len(indices), dim=0)])
        noise_pred = self.unet(latent_model_input, t, encoder_hidden_states=text_embed_input)['sample']
        _, noise_pred_uncond, noise_pred_cond = noise_pred.chunk(3)
        noise_pred = noise_pred_uncond + self.config["guidance_scale"] * (noise_pred_cond - noise_pred_uncond)
        return self.scheduler.step(noise_pred, t, x)['prev_sample']

    def _autocast(self):
        if self.device.type == "cuda" and self.config.get("autocast", True):
            return torch.autocast(device_type="cuda", dtype=torch.float16)    # run_tokenflow_pnp.py:220
        return contextlib.nullcontext()

    def draw_keyframes(self, n: int) -> torch.Tensor:
        """run_tokenflow_pnp.py:224 — one uniformly random frame inside every batch (CPU RNG)."""
        batch_size = self.config["batch_size"]
        return torch.randint(batch_size, (n // batch_size,)) + torch.arange(0, n, batch_size)

    def batched_denoise_step(self, x, t, indices):
        """run_tokenflow_pnp.py:220-233 (one process), or its frame-sharded form (world_size > 1).
        With config["fused_pass"] the pivotal samples and the frame samples go through the UNet in ONE
        call (the keyframe caches a block fills from the first part of the batch are consumed by the
        second part inside the same block) — identical arithmetic, half the kernel launches."""
        if self.config.get("fused_pass", False):
            with self._autocast():
                return self._fused_step(x, t, indices)
        if self.world_size > 1:
            with self._autocast():
                return self._sharded_step(x, t, indices)
        h = self.hooks
        batch_size = self.config["batch_size"]
        with self._autocast():
            pivotal_idx = self.draw_keyframes(len(x))
            self.keyframe_log.append(pivotal_idx.tolist())
            h.register_pivotal(self, True)
            self.denoise_step(x[pivotal_idx], t, indices[pivotal_idx])
            h.register_pivotal(self, False)
            per_pass = int(self.config.get("frames_per_pass", batch_size))
            if per_pass == batch_size:                                   # the reference's schedule (:229-231)
                denoised = []
                for i, b in enumerate(range(0, len(x), batch_size)):
                    h.register_batch_idx(self, i)
                    denoised.append(self.denoise_step(x[b:b + batch_size], t, indices[b:b + batch_size]))
                return torch.cat(denoised)
            # same arithmetic, fewer and larger UNet passes: frames of several batches in one pass, each frame
            # carrying its own (keyframe, previous keyframe, weight) — the per-frame table the kernels take
            denoised = []
            for b in range(0, len(x), per_pass):
                frames = list(range(b, min(len(x), b + per_pass)))
                h.register_frame_table(self, *self.frame_table(frames))
                denoised.append(self.denoise_step(x[b:b + per_pass], t, indices[b:b + per_pass]))
            return torch.cat(denoised)

    # ------------------------------------------------------------------------------------
    # multi-GPU: one process per GPU, frames sharded, keyframe tensors all-gathered
    # ------------------------------------------------------------------------------------
    def frame_table(self, frames):
        """Per-frame (keyframe, previous keyframe, blend weight) for global frame ids — the reference's
        batch_idx arithmetic (tokenflow_utils.py:331-333, :375-383) evaluated per frame."""
        from .ops import blend_weights
        B = self.config["batch_size"]
        w = blend_weights(B)
        kf_a = [g // B for g in frames]
        kf_b = [(g // B) - 1 if g >= B else -1 for g in frames]
        return kf_a, kf_b, [w[g % B] for g in frames]

    @torch.no_grad()
    def _sharded_step(self, x, t, indices):
        import torch.distributed as dist
        h, G, r = self.hooks, self.world_size, self.rank
        N, B = len(x), self.config["batch_size"]
        K = N // B
        assert N % G == 0, "frames must divide evenly over the ranks"
        pivotal_idx = self.draw_keyframes(N)                  # same CPU seed on every rank -> same keyframes
        self.keyframe_log.append(pivotal_idx.tolist())
        src_all = self.source_latents_t(int(t))[indices].to(x.device, x.dtype)
        h.register_time(self, int(t))
        # ---- pivotal pass: this rank's m of the 3K (stream, keyframe) samples ----
        shard = h.PivotalShard(G, r, K, self.group)
        lat, emb = [], []
        for i in shard.slots:
            i = min(i, 3 * K - 1)                             # padding slots recompute the last sample
            s, f = divmod(i, K)
            frame = int(pivotal_idx[f])
            lat.append(src_all[frame] if s == 0 else x[frame])
            emb.append(self.pnp_guidance_embeds[0] if s == 0 else self.text_embeds[s - 1])
        h.register_shard(self, shard)
        h.register_pivotal(self, True)
        self.unet(torch.stack(lat), t, encoder_hidden_states=torch.stack(emb))
        h.register_pivotal(self, False)
        h.register_shard(self, None)
        # ---- frame pass: this rank's contiguous frames, per-frame keyframe table ----
        per = N // G
        frames = list(range(r * per, (r + 1) * per))
        h.register_frame_table(self, *self.frame_table(frames))
        xs = x[frames[0]:frames[-1] + 1]
        latent_model_input = torch.cat([src_all[frames[0]:frames[-1] + 1], xs, xs])
        text = torch.cat([self.pnp_guidance_embeds.repeat(per, 1, 1), torch.repeat_interleave(self.text_embeds, per, dim=0)])
        noise_pred = self.unet(latent_model_input, t, encoder_hidden_states=text)['sample']
        _, npu, npc = noise_pred.chunk(3)
        noise_pred = npu + self.config["guidance_scale"] * (npc - npu)
        x_local = self.scheduler.step(noise_pred, t, xs)['prev_sample'].contiguous()
        out = torch.empty_like(x)
        dist.all_gather_into_tensor(out, x_local, group=self.group)
        return out

    def _timestep_pair(self, t):
        """(host int, device scalar) of a timestep without reading the device when `t` is a host value."""
        t_int = t if isinstance(t, int) else int(t)
        t_dev = self._t_dev.get(t_int)
        if t_dev is None:
            t_dev = self._t_dev[t_int] = torch.tensor(t_int, device=self.device)
        return t_int, t_dev

    def step_index(self, x, i: int, indices=None):
        """Denoising step number `i` of the schedule, addressed by index so that no device value is read
        back on the host (the reference's loop passes a CUDA 0-dim timestep, which costs a stream
        synchronisation per use)."""
        if indices is None:
            indices = torch.arange(len(x))
        return self.batched_denoise_step(x, self._t_host[i % len(self._t_host)], indices)

    @torch.no_grad()
    def _fused_step(self, x, t, indices):
        """One UNet call per denoising step and rank: [pivotal samples | this rank's frames x 3 streams]."""
        h, G, r = self.hooks, self.world_size, self.rank
        N, B = len(x), self.config["batch_size"]
        K = N // B
        assert N % G == 0, "frames must divide evenly over the ranks"
        t_int, t = self._timestep_pair(t)
        pivotal_idx = self.draw_keyframes(N)
        kf_list = pivotal_idx.tolist()
        self.keyframe_log.append(kf_list)
        src_all = self.source_latents_t(t_int)
        if not (indices.device.type == "cpu" and indices.numel() == src_all.shape[0]
                and torch.equal(indices, torch.arange(indices.numel()))):
            src_all = src_all[indices]                        # (identity in the drivers: all frames, in order)
        src_all = src_all.to(x.device, x.dtype)
        h.register_time(self, t_int)
        if G == 1:                                            # the reference's pivotal batch: [src | uncond | cond] x K
            shard = None
            x_kf = torch.stack([x[j] for j in kf_list])       # host-side index list: no index tensor upload
            piv_lat = torch.cat([torch.stack([src_all[j] for j in kf_list]), x_kf, x_kf])
            piv_emb = torch.cat([self.pnp_guidance_embeds.repeat(K, 1, 1),
                                 torch.repeat_interleave(self.text_embeds, K, dim=0)])
        else:                                                 # this rank's m of the 3K (stream, keyframe) samples
            shard = h.PivotalShard(G, r, K, self.group)
            lat, emb = [], []
            for i in shard.slots:
                i = min(i, 3 * K - 1)                         # padding slots recompute the last sample
                s_, f_ = divmod(i, K)
                frame = int(pivotal_idx[f_])
                lat.append(src_all[frame] if s_ == 0 else x[frame])
                emb.append(self.pnp_guidance_embeds[0] if s_ == 0 else self.text_embeds[s_ - 1])
            piv_lat, piv_emb = torch.stack(lat), torch.stack(emb)
        per = N // G
        lo = r * per
        frames = list(range(lo, lo + per))
        xs, srcs = x[lo:lo + per], src_all[lo:lo + per]
        n_piv = piv_lat.shape[0]
        latent_model_input = torch.cat([piv_lat, srcs, xs, xs])
        text = torch.cat([piv_emb, self.pnp_guidance_embeds.repeat(per, 1, 1),
                          torch.repeat_interleave(self.text_embeds, per, dim=0)])
        h.register_pivotal(self, False)
        h.register_shard(self, shard)
        h.register_frame_table(self, *self.frame_table(frames))
        h.register_fused(self, n_piv)
        try:
            noise_pred = self.unet(latent_model_input, t, encoder_hidden_states=text)['sample'][n_piv:]
        finally:
            h.register_fused(self, 0)
            h.register_shard(self, None)
        _, npu, npc = noise_pred.chunk(3)
        noise_pred = npu + self.config["guidance_scale"] * (npc - npu)
        x_local = self.scheduler.step(noise_pred, t_int, xs)['prev_sample'].contiguous()
        if G == 1:
            return x_local
        import torch.distributed as dist
        out = torch.empty_like(x)
        dist.all_gather_into_tensor(out, x_local, group=self.group)
        return out

    # ------------------------------------------------------------------------------------
    # host-buffer entry point (bench `e2e`): latents live in pinned host memory
    # ------------------------------------------------------------------------------------
    def edit_step_host(self, x_host: torch.Tensor, src_host_t: torch.Tensor, t: int, out_host: torch.Tensor):
        """One denoising step with HOST latents: H2D of this step's noisy latents and source latents,
        the step, D2H of the denoised latents, stream-synchronised before returning."""
        x = x_host.to(self.device, non_blocking=True)
        self._src_override = (int(t), src_host_t.to(self.device, non_blocking=True))
        try:
            y = self.batched_denoise_step(x, int(t), torch.arange(len(x_host)))
        finally:
            self._src_override = None
        out_host.copy_(y, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return out_host

    def sample_loop(self, x, indices=None, on_step: Optional[Callable] = None):
        """run_tokenflow_pnp.py:264-273 without the VAE decode."""
        if indices is None:
            indices = torch.arange(len(x))
        for i, t in enumerate(self._t_host):          # host ints: nothing is read back from the device per step
            x = self.batched_denoise_step(x, t, indices)
            if on_step is not None:
                on_step(i, t, x)
        return x


# --------------------------------------------------------------------------------------------
# synthetic inputs: no SD weights / VAE / CLIP exist here
# --------------------------------------------------------------------------------------------
def synthetic_inputs(n_frames: int, latent_size: int, ctx_dim: int, n_timesteps: int, seed: int = 1,
                     device="cpu", dtype=torch.float32, ctx_len: int = 77):
    """x ~ N(0,1) [N,4,L,L]; one source latent tensor per sampling timestep; text embeddings
    ~ N(0,1).  Deterministic in `seed` and independent of device."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_frames, 4, latent_size, latent_size, generator=g)
    text = torch.randn(2, ctx_len, ctx_dim, generator=g)
    pnp = torch.randn(1, ctx_len, ctx_dim, generator=g)
    ratio = 1000 // n_timesteps
    timesteps = [(n_timesteps - 1 - i) * ratio + 1 for i in range(n_timesteps)]
    src = {t: torch.randn(n_frames, 4, latent_size, latent_size, generator=g) for t in timesteps}
    conv = lambda z: z.to(device=device, dtype=dtype)
    return conv(x), conv(text), conv(pnp), {t: conv(v) for t, v in src.items()}


def write_latents_dir(path: str, src: Dict[int, torch.Tensor], prompt: str = "synthetic") -> str:
    """The preprocess -> edit hand-off format (preprocess.py:227-229, :313-314):
    <path>/latents/noisy_latents_<t>.pt + <path>/inversion_prompt.txt."""
    lat = os.path.join(path, "latents")
    os.makedirs(lat, exist_ok=True)
    for t, v in src.items():
        torch.save(v, os.path.join(lat, f"noisy_latents_{t}.pt"))
    with open(os.path.join(path, "inversion_prompt.txt"), "w") as f:
        f.write(prompt)
    return lat
