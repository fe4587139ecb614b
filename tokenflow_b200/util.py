"""Drop-in for the names the reference drivers and hooks import from `util`
(run_tokenflow_pnp.py:17 `save_video, seed_everything`; tokenflow_utils.py:5 `isinstance_str,
batch_cosine_sim`).  Video/image file I/O of reference util.py is host-side and out of scope
(SURVEY.md §2); `save_video` is kept so the drivers import and run.
"""
from __future__ import annotations

import random

import numpy as np
import torch


def isinstance_str(x: object, cls_name: str) -> bool:
    """True if any class in x's MRO is *named* cls_name (reference util.py:46-58): the hooks patch
    modules without importing their classes."""
    return any(c.__name__ == cls_name for c in type(x).__mro__)


def batch_cosine_sim(x, y):
    """Reference util.py:61-69.  Kept for API parity only: the TokenFlow block never materialises
    this matrix (tf_nn_field fuses normalise -> GEMM -> argmax); callers that really want the
    full matrix get the fp16 unit rows from the CUDA kernel and one library GEMM."""
    from . import tokenflow_utils as _tf

    if isinstance(x, list):
        x = torch.cat(x, dim=0)
    if isinstance(y, list):
        y = torch.cat(y, dim=0)
    ops = _tf._ops()
    return ops.unit_rows(x) @ ops.unit_rows(y).T


def seed_everything(seed: int):
    """Reference util.py:99-103."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    random.seed(seed)
    np.random.seed(seed)


def save_video(raw_frames: torch.Tensor, save_path: str, fps: int = 10):
    """Reference util.py:88-96 writes h264 through torchvision.io.write_video, which torchvision
    0.26 no longer ships; frames ([N,3,H,W] in [0,1]) are written with OpenCV instead."""
    import cv2

    frames = (raw_frames * 255).to(torch.uint8).cpu().permute(0, 2, 3, 1).numpy()
    h, w = frames.shape[1:3]
    writer = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
    try:
        for fr in frames:
            writer.write(cv2.cvtColor(fr, cv2.COLOR_RGB2BGR))
    finally:
        writer.release()
