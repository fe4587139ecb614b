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
        self._t_index = {t: i for i, t in enumerate(self._t_host)}
        # keyframe draws: world_size == 1 follows the reference (global CPU RNG, run_tokenflow_pnp.py:224); with
        # several ranks every rank must draw the SAME keyframes, so the draws come from a dedicated generator
        # seeded identically everywhere (config "keyframe_seed", default = the reference's seed 1) and are
        # independent of whatever else consumes the global RNG on a rank.  config["check_keyframes"] additionally
        # all-gathers the draw and asserts equality (debug / verify runs).
        self._kf_gen = None
        if world_size > 1 or "keyframe_seed" in self.config:
            self._kf_gen = torch.Generator().manual_seed(int(self.config.get("keyframe_seed", self.config.get("seed", 1))))
        self.comm = None                      # ops.Communicator (C-ABI NCCL all-gather), see attach_communicator
        # DDIM coefficients per schedule position, on the device, for the fused CFG+DDIM kernel:
        # sqrt(1-a_t), 1/sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev) as the fp32 values the eager expression uses
        self._coef_table = self._make_coef_table()
        self._graphs = {}                     # injection variant -> captured step
        self._graph_pool = None
        self._g_static = None
        self._text_cache = {}
        self._shard_cache = {}
        self._side_stream = None

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
        if self._kf_gen is None:
            r = torch.randint(batch_size, (n // batch_size,))
        else:
            r = torch.randint(batch_size, (n // batch_size,), generator=self._kf_gen)
        idx = r + torch.arange(0, n, batch_size)
        if self.world_size > 1 and self.config.get("check_keyframes", False):
            import torch.distributed as dist
            mine = idx.to(self.device if dist.get_backend(self.group) == "nccl" else "cpu")
            everyone = [torch.empty_like(mine) for _ in range(self.world_size)]
            dist.all_gather(everyone, mine, group=self.group)
            for r_, other in enumerate(everyone):
                if not torch.equal(other.cpu(), idx):
                    raise RuntimeError(f"rank {self.rank} drew keyframes {idx.tolist()} but rank {r_} drew "
                                       f"{other.cpu().tolist()}: the ranks' keyframe generators are out of step")
        return idx

    def _make_coef_table(self):
        import numpy as np
        sch, rows = self.scheduler, []
        ratio = sch.num_train_timesteps // sch.num_inference_steps
        for t in self._t_host:
            a_t = float(sch._alpha(t))
            a_prev = float(sch._alpha(t - ratio))
            s1, s2 = np.float32((1 - a_t) ** 0.5), np.float32(a_t ** 0.5)
            rows.append([float(s1), float(np.float32(1.0) / s2), float(np.float32(a_prev ** 0.5)),
                         float(np.float32((1 - a_prev) ** 0.5))])
        return torch.tensor(rows, dtype=torch.float32, device=self.device)

    def attach_communicator(self, comm):
        """Route the pivotal pass's all-gathers through the C ABI (tf_allgather) instead of torch.distributed."""
        self.comm = comm

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
        shard = h.PivotalShard(G, r, K, self.group, comm=self.comm, token_split=self.config.get("token_split", True))
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

    # ------------------------------------------------------------------------------------
    # fused step: ONE UNet call per denoising step and rank, [pivotal samples | this rank's frames x 3 streams]
    # ------------------------------------------------------------------------------------
    def _pivotal_slots(self, K: int):
        """(stream, keyframe) of every pivotal sample this rank runs, in batch order.  One rank: the reference's
        [src | uncond | cond] x K batch.  Several ranks: this rank's m = ceil(3K/G) slots of that order (the tail is
        padded with repeats of the last sample, whose results are ignored)."""
        G, r = self.world_size, self.rank
        if G == 1:
            return [divmod(i, K) for i in range(3 * K)], None
        key = (K, id(self.comm))
        shard = self._shard_cache.get(key)
        if shard is None:                     # one shard context per (K, communicator): it caches device index tensors
            shard = self._shard_cache[key] = self.hooks.PivotalShard(G, r, K, self.group, comm=self.comm,
                                                                     token_split=self.config.get("token_split", True))
        return [divmod(min(i, 3 * K - 1), K) for i in shard.slots], shard

    def _fused_text(self, slots, per):
        key = (tuple(slots), per)
        text = self._text_cache.get(key)
        if text is None:
            emb = [self.pnp_guidance_embeds[0] if s_ == 0 else self.text_embeds[s_ - 1] for s_, _ in slots]
            text = torch.cat([torch.stack(emb), self.pnp_guidance_embeds.repeat(per, 1, 1),
                              torch.repeat_interleave(self.text_embeds, per, dim=0)])
            self._text_cache = {key: text}
        return text

    def _fused_compute(self, x, src_all, piv_idx, t_dev, t_int, coef, slots, shard):
        """Device work of one fused step.  Everything that varies from step to step arrives in device tensors
        (`piv_idx`: which latents are the pivotal samples, `t_dev`, `coef`: the DDIM coefficients), so the same
        function body can be captured once into a CUDA graph and replayed (run_tokenflow_pnp.py:195-233)."""
        h, G, r = self.hooks, self.world_size, self.rank
        N = x.shape[0]
        per = N // G
        lo = r * per
        n_piv = len(slots)
        piv_lat = torch.cat([src_all, x]).index_select(0, piv_idx)       # slot (s, f): src[kf_f] if s == 0 else x[kf_f]
        xs, srcs = x[lo:lo + per], src_all[lo:lo + per]
        latent_model_input = torch.cat([piv_lat, srcs, xs, xs])
        text = self._fused_text(slots, per)
        h.register_time(self, t_int)
        h.register_pivotal(self, False)
        h.register_shard(self, shard)
        h.register_frame_table(self, *self.frame_table(list(range(lo, lo + per))))
        h.register_fused(self, n_piv)
        try:
            noise_pred = self.unet(latent_model_input, t_dev, encoder_hidden_states=text)['sample'][n_piv:]
        finally:
            h.register_fused(self, 0)
            h.register_shard(self, None)
        _, npu, npc = noise_pred.chunk(3)
        ops = self._cuda_ops()
        if ops is not None and coef is not None and npu.dtype == torch.float16 and xs.dtype == torch.float16:
            x_local = ops.cfg_ddim(npu, npc, xs, coef, self.config["guidance_scale"])     # one kernel, same roundings
        else:
            noise_pred = npu + self.config["guidance_scale"] * (npc - npu)
            x_local = self.scheduler.step(noise_pred, t_int, xs)['prev_sample'].contiguous()
        if G == 1:
            return x_local
        if self.comm is not None and x_local.is_cuda:
            return self.comm.all_gather(x_local)
        import torch.distributed as dist
        out = torch.empty_like(x)
        dist.all_gather_into_tensor(out, x_local, group=self.group)
        return out

    def _dual_compute(self, x, src_all, piv_idx, t_dev, t_int, coef, slots, shard):
        """The same step as `_fused_compute` as TWO UNet calls on two CUDA streams: the pivotal samples (few, and all
        the collectives when sharded) on a side stream, this rank's frames on the current stream; each TokenFlow block
        of the frame call waits for the event its pivotal counterpart recorded (`register_dual_stream`).  The pivotal
        chain — latency-bound small kernels plus the all-gathers — then runs under the frame chain's compute."""
        h, G, r = self.hooks, self.world_size, self.rank
        N = x.shape[0]
        per = N // G
        lo = r * per
        main = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        side = self._side_stream
        piv_lat = torch.cat([src_all, x]).index_select(0, piv_idx)
        text = self._fused_text(slots, per)
        n_piv = len(slots)
        piv_emb, frame_emb = text[:n_piv], text[n_piv:]
        xs, srcs = x[lo:lo + per], src_all[lo:lo + per]
        frame_in = torch.cat([srcs, xs, xs])
        h.register_time(self, t_int)
        h.register_fused(self, 0)
        h.register_dual_stream(self, True)
        try:
            side.wait_stream(main)
            with torch.cuda.stream(side):                    # ---- pivotal chain
                h.register_pivotal(self, True)
                h.register_shard(self, shard)
                self.unet(piv_lat, t_dev, encoder_hidden_states=piv_emb)
            h.register_pivotal(self, False)                  # ---- frame chain
            h.register_shard(self, None)
            h.register_frame_table(self, *self.frame_table(list(range(lo, lo + per))))
            noise_pred = self.unet(frame_in, t_dev, encoder_hidden_states=frame_emb)['sample']
            main.wait_stream(side)                           # join (piv_lat and the caches stay referenced until here)
        finally:
            h.register_dual_stream(self, False)
            h.register_shard(self, None)
        del piv_lat
        _, npu, npc = noise_pred.chunk(3)
        ops = self._cuda_ops()
        if ops is not None and coef is not None and npu.dtype == torch.float16 and xs.dtype == torch.float16:
            x_local = ops.cfg_ddim(npu, npc, xs, coef, self.config["guidance_scale"])
        else:
            noise_pred = npu + self.config["guidance_scale"] * (npc - npu)
            x_local = self.scheduler.step(noise_pred, t_int, xs)['prev_sample'].contiguous()
        if G == 1:
            return x_local
        if self.comm is not None and x_local.is_cuda:
            return self.comm.all_gather(x_local)
        import torch.distributed as dist
        out = torch.empty_like(x)
        dist.all_gather_into_tensor(out, x_local, group=self.group)
        return out

    def _step_compute(self, *a):
        # Off unless asked for: on one GPU the concurrent chains slow each other down (3.32 vs 3.50 frames/s at C2), and
        # with real NCCL ranks the 8-GPU run of this schedule did not finish (profiles/README.md) — only its
        # single-process forms are verified (tests/test_gpu_round2.py).
        dual = bool(self.config.get("dual_stream", False))
        if dual and self.device.type == "cuda":
            return self._dual_compute(*a)
        return self._fused_compute(*a)

    def _cuda_ops(self):
        """The CUDA op object if the hooks run on it (None under the oracle test seam / on CPU)."""
        if self.device.type != "cuda":
            return None
        ops = self.hooks._ops() if hasattr(self.hooks, "_ops") else None
        return ops if hasattr(ops, "cfg_ddim") else None

    def _piv_index_list(self, kf_list, slots, N):
        return [kf_list[f_] if s_ == 0 else N + kf_list[f_] for s_, f_ in slots]

    def _variant(self, t_int):
        """Which hooks inject at this timestep — the only way `t` changes the captured kernel sequence."""
        if self.config.get("mode", "pnp") != "pnp":
            return (False, False)
        qk = {int(v) for v in (self.qk_injection_timesteps.tolist() if torch.is_tensor(self.qk_injection_timesteps)
                               else self.qk_injection_timesteps)}
        conv = {int(v) for v in (self.conv_injection_timesteps.tolist() if torch.is_tensor(self.conv_injection_timesteps)
                                 else self.conv_injection_timesteps)}
        return (t_int in qk or t_int == 1000, t_int in conv or t_int == 1000)

    @torch.no_grad()
    def _fused_step(self, x, t, indices):
        """One UNet call per denoising step and rank: [pivotal samples | this rank's frames x 3 streams]."""
        N, B = len(x), self.config["batch_size"]
        K = N // B
        assert N % self.world_size == 0, "frames must divide evenly over the ranks"
        t_int, t_dev = self._timestep_pair(t)
        pivotal_idx = self.draw_keyframes(N)
        kf_list = pivotal_idx.tolist()
        self.keyframe_log.append(kf_list)
        src_all = self.source_latents_t(t_int)
        if not (indices.device.type == "cpu" and indices.numel() == src_all.shape[0]
                and torch.equal(indices, torch.arange(indices.numel()))):
            src_all = src_all[indices]                        # (identity in the drivers: all frames, in order)
        src_all = src_all.to(x.device, x.dtype)
        slots, shard = self._pivotal_slots(K)
        idx_host = torch.tensor(self._piv_index_list(kf_list, slots, N), dtype=torch.int64)
        i = self._t_index.get(t_int)
        coef = self._coef_table[i] if i is not None else None
        if self.config.get("cuda_graph", False) and self.device.type == "cuda" and coef is not None:
            return self._graph_replay(x, src_all, idx_host, t_int, i, slots, shard)
        piv_idx = idx_host.to(x.device, non_blocking=True) if x.is_cuda else idx_host
        return self._step_compute(x, src_all, piv_idx, t_dev, t_int, coef, slots, shard)

    # ------------------------------------------------------------------------------------
    # CUDA graphs: the fused step's shape is static, so it is captured once per injection
    # variant (PnP: q/k + conv injection, conv injection only, none) and replayed.  Per replay the host only
    # refreshes five small static inputs: latents, source latents, the pivotal gather index, the timestep and
    # the DDIM coefficients.  TMA descriptors, frame tables and attention tables are kernel parameters baked at
    # capture; NCCL all-gathers are captured in-graph.
    # ------------------------------------------------------------------------------------
    def _graph_replay(self, x, src_all, idx_host, t_int, i, slots, shard):
        variant = self._variant(t_int)
        st = self._g_static
        if st is None or st["x"].shape != x.shape or st["x"].dtype != x.dtype:
            st = self._g_static = {
                "x": torch.empty_like(x), "src": torch.empty_like(x),
                "idx": torch.zeros(len(idx_host), dtype=torch.int64, device=x.device),
                "t": torch.zeros((), dtype=torch.int64, device=x.device),
                "coef": torch.zeros(4, dtype=torch.float32, device=x.device)}
            self._graphs = {}
        if x.data_ptr() != st["x"].data_ptr():
            st["x"].copy_(x, non_blocking=True)
        if src_all.data_ptr() != st["src"].data_ptr():
            st["src"].copy_(src_all, non_blocking=True)
        st["idx"].copy_(idx_host.pin_memory(), non_blocking=True)
        st["t"].copy_(self._t_dev[t_int], non_blocking=True)
        st["coef"].copy_(self._coef_table[i], non_blocking=True)
        entry = self._graphs.get(variant)
        if entry is None:
            entry = self._graphs[variant] = self._capture(st, t_int, slots, shard)
        entry["graph"].replay()
        entry["replays"] += 1
        return entry["out"].clone()

    def _capture(self, st, t_int, slots, shard):
        ops = self._cuda_ops()
        run = lambda: self._step_compute(st["x"], st["src"], st["idx"], st["t"], t_int, st["coef"], slots, shard)
        # warm-up on a side stream (cuDNN autotuning, lazy initialisation, allocator growth) — not captured
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved_timing = ops._timing if ops is not None else None
        if ops is not None:
            ops._timing = None
        with torch.cuda.stream(side):
            for _ in range(int(self.config.get("graph_warmup", 2))):
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        events = [] if saved_timing is not None else None
        launches0 = ops.launch_count() if ops is not None else 0
        if ops is not None:
            ops._timing = events                # per-launch EXTERNAL events recorded as graph nodes
        try:
            with torch.cuda.graph(graph, pool=self._graph_pool, capture_error_mode="thread_local"):
                out = run()
        finally:
            if ops is not None:
                ops._timing = saved_timing
        if self._graph_pool is None:
            self._graph_pool = graph.pool()
        return {"graph": graph, "out": out, "events": events, "replays": 0,
                "launches": (ops.launch_count() - launches0) if ops is not None else 0}

    def graph_launches_per_step(self) -> int:
        """Kernels of this library inside one replay of the most recently used step graph."""
        used = [e for e in self._graphs.values() if e["replays"]]
        return max((e["launches"] for e in used), default=0)

    def mark_graph_replays(self):
        """Start of a measured region: `graph_kernel_times(since_mark=True)` then covers the replays after this call."""
        for entry in self._graphs.values():
            entry["mark"] = entry["replays"]

    @staticmethod
    def aggregate_graph_events(entries, since_mark=False):
        """{kernel: {"launches", "ms", "work"}} over the replays of the captured step graphs.  The event nodes are part
        of a graph and every replay overwrites their timestamps, so what can be read is the LAST replay of each
        variant; a variant that was replayed n times contributes n times its last replay.  Returns (totals, steps)."""
        agg, steps = {}, 0
        for entry in entries:
            n = entry["replays"] - (entry.get("mark", 0) if since_mark else 0)
            if n <= 0 or not entry.get("events"):
                continue
            steps += n
            for name, work, s_, e_ in entry["events"]:
                a = agg.setdefault(name, {"launches": 0, "ms": 0.0, "work": 0.0})
                a["launches"] += n
                a["ms"] += n * s_.elapsed_time(e_)
                a["work"] += n * work
        return agg, steps

    def graph_kernel_times(self, since_mark=False):
        """Per-kernel totals of the step graphs' per-launch event nodes (see `aggregate_graph_events`)."""
        torch.cuda.synchronize()
        return self.aggregate_graph_events(self._graphs.values(), since_mark)

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
