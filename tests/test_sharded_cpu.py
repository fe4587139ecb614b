"""CPU tier: the multi-GPU path (frames sharded, keyframe tensors all-gathered) on 2 gloo ranks with
the oracle ops == the single-process loop.  Covers the shard plan, the all-gather ordering, the
per-frame keyframe/weight tables and the sharded attention table."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tokenflow_b200 import tokenflow_utils as tfu


def _init_file(tmp_path_factory=None):
    """file:// rendezvous: no port to race for."""
    import tempfile
    fd, path = tempfile.mkstemp(prefix="tf_b200_rdzv_")
    os.close(fd)
    os.unlink(path)
    return path


def _edit(world, rank, mode, steps, fused=False):
    from oracle.oracle_ops import OracleOps
    from tokenflow_b200 import sd_unet
    from tokenflow_b200.editor import TokenFlowEditor, synthetic_inputs
    from tokenflow_b200.scheduler import DDIMScheduler
    tfu._install_ops_for_testing(OracleOps())
    unet = sd_unet.build_unet("tiny", seed=1)
    cfg = {"n_frames": 8, "batch_size": 2, "n_timesteps": steps, "guidance_scale": 7.5, "mode": mode,
           "pnp_attn_t": 0.5, "pnp_f_t": 0.8, "start": 0.9, "fused_pass": fused}
    x, text, pnp, src = synthetic_inputs(8, 16, unet.config.cross_attention_dim, steps, seed=1, ctx_len=7)
    ed = TokenFlowEditor(unet, DDIMScheduler(), tfu, cfg, text, pnp, source_latents=lambda t: src[t],
                         world_size=world, rank=rank)
    ed.init_method()
    torch.manual_seed(1)
    return ed.sample_loop(x), ed.keyframe_log


def _worker(rank, world, rdzv, mode, steps, q, fused=False):
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method=f"file://{rdzv}", rank=rank, world_size=world)
    try:
        out, kf = _edit(world, rank, mode, steps, fused)
        # plain Python data on the queue: a torch tensor would travel by file-descriptor passing, which needs
        # the sender alive until the parent has rebuilt it (the worker exits right after the put)
        q.put((rank, out.numpy().tolist(), kf))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,steps,fused", [("pnp", 2, False), ("sdedit", 10, False), ("pnp", 2, True),
                                              ("sdedit", 10, True)])
def test_two_rank_edit_equals_single_process(mode, steps, fused):
    want, kf_want = _edit(1, 0, mode, steps)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    rdzv = _init_file()
    procs = [ctx.Process(target=_worker, args=(r, 2, rdzv, mode, steps, q, fused)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, kf in results:
        assert kf == kf_want
        assert torch.allclose(torch.tensor(out), want, atol=2e-4, rtol=1e-4), f"rank {rank}"


def test_shard_plan_and_attention_table():
    K = 5
    for G in (2, 4, 8):
        seen = []
        for r in range(G):
            sh = tfu.PivotalShard(G, r, K)
            assert len(sh.slots) == -(-15 // G)
            seen += sh.slots
            tab = sh.attention_table(False)
            tab_inj = sh.attention_table(True)
            for j, i in enumerate(sh.slots):
                if i >= 15:
                    assert tab[j][3] == 1 and tab_inj[j][3] == 1
                    continue
                s, f = divmod(i, K)
                if s == 0:
                    assert tab[j] == (j, i, i, 1) and tab_inj[j] == (i, i, i, 1)
                else:
                    assert tab[j] == (j, s * K, s * K, K)
                    assert tab_inj[j] == (f, 0, s * K, K)       # q and k of the source stream, own v
        assert seen[:15] == list(range(15)) and len(seen) == G * -(-15 // G)


def test_frame_table_matches_batch_idx_arithmetic():
    from oracle import tokenflow_oracle as O
    from tokenflow_b200.editor import TokenFlowEditor
    ed = TokenFlowEditor.__new__(TokenFlowEditor)
    ed.config = {"batch_size": 8}
    kf_a, kf_b, w = TokenFlowEditor.frame_table(ed, list(range(5, 10)))     # rank 1 of 8 at N=40: spans batches 0,1
    assert kf_a == [0, 0, 0, 1, 1] and kf_b == [-1, -1, -1, 0, 0]
    ref = O.blend_weights(1, 8)
    assert abs(w[3] - float(ref[0])) < 1e-7 and abs(w[4] - float(ref[1])) < 1e-7
