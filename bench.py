#!/usr/bin/env python
"""bench.py — frames/s of the 40-frame 512x512 SD1.5 50-step TokenFlow PnP edit (BASELINE.json
configs[1], "C2"): N=40 frames, keyframe stride B=8 -> K=5 keyframes per step, random-init
SD1.5-shape UNet in fp16, synthetic latents (no SD weights / VAE / CLIP exist offline).

A "step" is one denoising step of the edit = the pivotal pass over the K keyframes (extended
attention, caches filled) + the N/B frame passes (NN field + propagation) + CFG + DDIM update.
frames/s = N / (50 * mean step time): --steps K times K consecutive denoising steps of the 50.

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA kernels, sm_100a)
  python bench.py --impl reference ...                       the reference's algorithm on host cores

One JSON line on stdout (rank 0).  Keys follow the driver contract; `roofline` describes the dominant
hot-path kernel (time measured live with CUDA events around every launch inside the timed region),
`cpu_baseline` the oracle port timed on the host cores on a bounded sample, `e2e` the same metric
through the public editor call with pinned HOST latents (H2D + D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_FRAMES, BATCH, N_TIMESTEPS, LATENT = 40, 8, 50, 64
METRIC = "frames/sec for 40-frame 512x512 SD1.5 50-step edit"
WORKLOAD = "C2: 40-frame 512x512 SD1.5 PnP 50-step edit, B=8 (K=5 keyframes), random-init UNet fp16, synthetic latents"


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def build_editor(device, world=1, rank=0, seed=1, channels_last=True, frames_per_pass=BATCH, fused_pass=True):
    from tokenflow_b200 import sd_unet, tokenflow_utils as tfu
    from tokenflow_b200.editor import TokenFlowEditor, synthetic_inputs
    from tokenflow_b200.scheduler import DDIMScheduler
    # multi-GPU: draw the weights on the device (torchrun pins OMP_NUM_THREADS=1 and a CPU init of the 860M
    # parameters then takes minutes per rank); single GPU keeps the device-independent CPU init
    unet = sd_unet.build_unet("sd15", seed=seed, device=device, dtype=torch.float16, init_on_device=world > 1)
    if channels_last:
        unet = unet.to(memory_format=torch.channels_last)
    cfg = {"n_frames": N_FRAMES, "batch_size": BATCH, "n_timesteps": N_TIMESTEPS, "guidance_scale": 7.5,
           "mode": "pnp", "pnp_attn_t": 0.5, "pnp_f_t": 0.8, "frames_per_pass": frames_per_pass,
           "fused_pass": bool(fused_pass)}
    x, text, pnp, src = synthetic_inputs(N_FRAMES, LATENT, unet.config.cross_attention_dim, N_TIMESTEPS, seed=seed,
                                         device=device, dtype=torch.float16)
    ed = TokenFlowEditor(unet, DDIMScheduler(), tfu, cfg, text, pnp, source_latents=lambda t: src[t],
                         world_size=world, rank=rank)
    ed.init_method()
    return ed, x, src


def run_ours(args):
    rank, local_rank, world = dist_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = os.environ.get("TF_BENCH_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)
    from tokenflow_b200 import tokenflow_utils as tfu
    ops = tfu._ops()                                     # CudaOps: raises if the .so / B200 is missing
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    ed, x0, src = build_editor(device, world, rank, channels_last=not args.no_channels_last,
                               frames_per_pass=args.frames_per_pass, fused_pass=bool(args.fused_pass))
    timesteps = [int(t) for t in ed.scheduler.timesteps]
    indices = torch.arange(N_FRAMES)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_device(x, i):
        return ed.step_index(x, i, indices)          # by schedule index: no device read-back of the timestep

    # ---- device-resident measurement (`value`) ----
    torch.manual_seed(1)
    x = x0.clone()
    for i in range(args.warmup):
        x = step_device(x, i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ops.enable_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        x = step_device(x, args.warmup + i)
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ops.launch_count() - launches0
    kernel_times = ops.timing_summary()
    ops.enable_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    finite = bool(torch.isfinite(x.float()).all().item())

    # ---- end-to-end through the public call with pinned host latents (`e2e`) ----
    ms_e2e = float("nan")
    if not args.skip_e2e:
        x_host = x0.cpu().pin_memory()
        src_host = {t: v.cpu().pin_memory() for t, v in src.items()}
        out_host = torch.empty_like(x_host).pin_memory()
        torch.manual_seed(1)
        for i in range(min(args.warmup, 3)):
            ed.edit_step_host(x_host, src_host[timesteps[i % N_TIMESTEPS]], timesteps[i % N_TIMESTEPS], out_host)
            x_host.copy_(out_host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            t = timesteps[(args.warmup + i) % N_TIMESTEPS]
            ed.edit_step_host(x_host, src_host[t], t, out_host)
            x_host.copy_(out_host)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)

    # max over ranks
    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], device=device, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms_total, ms_e2e = tt.tolist()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    peaks = measured_peaks()
    ms_step = ms_total / args.steps
    fps = N_FRAMES / (N_TIMESTEPS * ms_step / 1e3)
    if ms_e2e == ms_e2e:                                  # not NaN: the host-buffer leg ran
        ms_step_e2e = round(ms_e2e / args.steps, 3)
        fps_e2e = round(N_FRAMES / (N_TIMESTEPS * ms_step_e2e / 1e3), 4)
    else:                                                 # --skip-e2e (profiling runs)
        ms_step_e2e = fps_e2e = None
    lat_bytes = x0.numel() * x0.element_size()

    # dominant hot-path kernel by summed launch time inside the timed region
    roofline = None
    if kernel_times:
        dom = max(kernel_times, key=lambda k_: kernel_times[k_]["ms"])
        kt = kernel_times[dom]
        traffic = None
        tpath = os.path.join(REPO, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(dom)
        if dom in ("tf_ext_attn", "tf_nn_field"):
            achieved = kt["work"] / (kt["ms"] * 1e-3) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["tf_sustained"],
                        "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_sustained"], 4), "traffic": traffic,
                        "peak_source": f"{peaks['source']} sustained cuBLAS bf16 (kernel timed inside a long step)"}
        else:
            achieved = kt["work"] / (kt["ms"] * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"],
                        "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic,
                        "peak_source": f"{peaks['source']} HBM copy"}
        roofline["launches"] = kt["launches"]
        roofline["avg_launch_ms"] = round(kt["ms"] / kt["launches"], 4)
    per_kernel = {}
    for name, kt in kernel_times.items():
        rate = kt["work"] / (kt["ms"] * 1e-3)
        per_kernel[name] = {"launches": kt["launches"], "ms_per_step": round(kt["ms"] / args.steps, 3),
                            ("tflops" if name in ("tf_ext_attn", "tf_nn_field") else "gbs"):
                                round(rate / (1e12 if name in ("tf_ext_attn", "tf_nn_field") else 1e9), 2)}

    cpu = cpu_baseline_sample() if (world == 1 and not args.no_cpu_baseline) else None

    line = {
        "metric": METRIC, "value": round(fps, 4), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_frames": N_FRAMES, "keyframes": N_FRAMES // BATCH, "ddim_steps": N_TIMESTEPS,
                   "frames_per_sec_definition": "n_frames / (50 * mean denoising-step time over the timed steps)",
                   "parallelism": f"frames sharded over {world} GPU(s)" if world > 1 else "single GPU",
                   "frames_per_pass": (N_FRAMES // world) if (world > 1 or args.fused_pass) else args.frames_per_pass,
                   "unet_calls_per_step": 1 if args.fused_pass else (2 if world > 1 else 1 + -(-N_FRAMES // args.frames_per_pass)),
                   "l2": "inputs > L2: every step streams ~10 GB of activations through the UNet (no flush needed)"},
        "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": ms_step_e2e,
                "h2d_bytes_per_step": 2 * lat_bytes, "d2h_bytes_per_step": lat_bytes},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "kernels": per_kernel,
        "hot_path_ms_per_step": round(sum(k_["ms"] for k_ in kernel_times.values()) / args.steps, 3),
        "finite": finite,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
class CpuSampler:
    """Bounded samples of the C2 denoising step on the host cores (fp32, all threads) — the
    reference's algorithm through the oracle port.

    The full step is ~20 minutes on 8 cores (135 UNet sample-forwards + ~11 TFLOP of hot path), so
    each sample measures the step's pieces at full 512x512 resolution and composes them with the
    exact op counts of C2:
      body   : one frame pass of ONE frame (3 stream samples) through the hooked SD1.5-shape UNet with
               the oracle ops; per-sample body time x 3*(K+N) sample-forwards per step
      nn/prop: the NN field of that frame against one keyframe and its propagation, timed inside the
               body pass, x the C2 counts (2N-B frame/keyframe pairs, N frames)
      attn   : the oracle's extended attention for ONE head and ONE keyframe's queries against the K*S
               keys of an extended stream, at each of the 4 UNet levels; x heads x (2K+1) x blocks
               (uncond+cond: K query frames x K*S keys each; source: K frames x S keys = 1 such unit)"""

    SD15_LEVELS = ((4096, 320, 8, 5), (1024, 640, 8, 5), (256, 1280, 8, 5), (64, 1280, 8, 1))   # (S, dim, heads, blocks)

    def __init__(self, kind="sd15", latent=LATENT, ctx_dim=768, levels=None):
        from oracle.oracle_ops import OracleOps
        from tokenflow_b200 import sd_unet, tokenflow_utils as tfu
        from tokenflow_b200.editor import TokenFlowEditor, synthetic_inputs
        from tokenflow_b200.scheduler import DDIMScheduler
        # all host cores, capped at 32: the oracle's eager PyTorch ops stop scaling (and then regress from
        # thread oversubscription) well before that on the many-core GPU hosts
        torch.set_num_threads(min(os.cpu_count() or 1, 32))
        self.tfu = tfu

        class TimedOracle(OracleOps):
            def __init__(self):
                self.t = {"nn": 0.0, "prop": 0.0}

            def nn_field(self, *a, **k):
                t0 = time.perf_counter(); r = super().nn_field(*a, **k); self.t["nn"] += time.perf_counter() - t0; return r

            def propagate(self, *a, **k):
                t0 = time.perf_counter(); r = super().propagate(*a, **k); self.t["prop"] += time.perf_counter() - t0; return r

        self.ops = TimedOracle()
        tfu._install_ops_for_testing(self.ops)
        with torch.no_grad():
            self.levels = levels or self.SD15_LEVELS
            unet = sd_unet.build_unet(kind, seed=1)
            cfg = {"n_frames": 1, "batch_size": 1, "n_timesteps": N_TIMESTEPS, "guidance_scale": 7.5, "mode": "pnp"}
            self.x, text, pnp, src = synthetic_inputs(1, latent, ctx_dim, N_TIMESTEPS, seed=1)
            self.ed = TokenFlowEditor(unet, DDIMScheduler(), tfu, cfg, text, pnp, source_latents=lambda t: src[t])
            self.ed.init_method()
            self.t0 = self.ed.scheduler.timesteps[0]
            tfu.register_pivotal(self.ed, True)
            self.ed.denoise_step(self.x, self.t0, torch.arange(1))      # fills the caches (K=1), untimed
            tfu.register_pivotal(self.ed, False)
            tfu.register_batch_idx(self.ed, 0)
        tfu._install_ops_for_testing(None)

    def step(self):
        from oracle import tokenflow_oracle as O
        K = N_FRAMES // BATCH
        self.tfu._install_ops_for_testing(self.ops)
        try:
            with torch.no_grad():
                self.ops.t = {"nn": 0.0, "prop": 0.0}
                t0 = time.perf_counter()
                self.ed.denoise_step(self.x, self.t0, torch.arange(1))  # 3 sample-forwards + NN(1 pair) + propagate(1 frame)
                t_pass = time.perf_counter() - t0
                t_nn_pair, t_prop_frame = self.ops.t["nn"], self.ops.t["prop"]
                t_body_sample = (t_pass - t_nn_pair - t_prop_frame) / 3.0
                t_attn = 0.0
                for S, dim, heads, blocks in self.levels:
                    d = dim // heads
                    q = torch.randn(1, S, d)
                    k, v = torch.randn(1, K * S, d), torch.randn(1, K * S, d)
                    t0 = time.perf_counter()
                    sim = torch.bmm(q, k.transpose(-1, -2)) * d ** -0.5        # oracle extended_attention, one unit
                    torch.bmm(sim.softmax(dim=-1), v)
                    t_attn += (time.perf_counter() - t0) * heads * (2 * K + 1) * blocks
        finally:
            self.tfu._install_ops_for_testing(None)
        t_step = (t_body_sample * 3 * (K + N_FRAMES) + t_attn + t_nn_pair * (2 * N_FRAMES - BATCH)
                  + t_prop_frame * N_FRAMES)
        desc = ("per step: one 512x512 frame pass (3 UNet sample-forwards + NN field vs 1 keyframe + propagation) and "
                "one (head, query-frame) unit of K=5 extended attention per UNet level, fp32, composed with the C2 op "
                f"counts (body {t_body_sample:.2f}s/sample x135, attn {t_attn:.1f}s, nn {t_nn_pair:.2f}s/pair x72, "
                f"prop {t_prop_frame:.3f}s/frame x40)")
        return t_step, desc


def cpu_baseline_sample():
    t_step, desc = CpuSampler().step()
    return {"value": round(N_FRAMES / (N_TIMESTEPS * t_step), 6), "unit": "frames/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": desc, "s_per_step_extrapolated": round(t_step, 2)}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    sampler = CpuSampler()
    times, desc = [], ""
    for i in range(args.warmup + args.steps):
        t_step, desc = sampler.step()
        if i >= args.warmup:
            times.append(t_step)
    t_step = sum(times) / len(times)
    fps = N_FRAMES / (N_TIMESTEPS * t_step)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": round(fps, 6), "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_step * 1e3, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference algorithm (oracle port) on host cores; each step is a bounded "
                       "sample of the C2 step composed with exact op counts"},
            "cpu_baseline": {"value": round(fps, 6), "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": round(fps, 6), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-channels-last", action="store_true", help="UNet body in NCHW instead of channels_last")
    ap.add_argument("--frames-per-pass", type=int, default=N_FRAMES,
                    help="frames per frame-pass UNet call (8 = the reference's per-batch schedule; default: all "
                         "frames of the GPU in one pass with per-frame keyframe tables — identical results)")
    ap.add_argument("--fused-pass", type=int, default=1,
                    help="1: one UNet call per step and GPU ([pivotal samples | frames], keyframe caches filled and "
                         "consumed inside each block); 0: the reference's pivotal pass + frame passes")
    ap.add_argument("--cudnn-benchmark", type=int, default=1, help="torch.backends.cudnn.benchmark for the UNet body convs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
