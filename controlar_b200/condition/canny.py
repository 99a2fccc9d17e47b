"""Canny control-map detector on the GPU — same surface as the reference's `condition/canny.py` (`CannyDetector()(img, low, high)`,
(H, W, 3) uint8 image in, (H, W) uint8 map of {0, 255} out), bit-exact against the `cv2.Canny` call it replaces
(condition/canny.py:14; integer arithmetic, csrc/frontend.cuh, oracle/canny_oracle.py).  No OpenCV, no host round trip of the image
when the input already lives on the GPU."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib
from .._lib import check, cur_stream, _ptr


def canny_cuda(img: torch.Tensor, low_threshold: float = 100, high_threshold: float = 200, sweeps_per_call: int = 8) -> torch.Tensor:
    """img: CUDA tensor (H, W, C) or (H, W), values 0..255 (any dtype; cast to uint8 like the reference's `.astype(np.uint8)`).
    Returns a CUDA uint8 tensor (H, W) with 255 on edges.  The hysteresis runs in sweeps on the device; only the "did the last sweep
    still change something" flag is read by the host (the reference's call is synchronous host code anyway)."""
    if img.device.type != "cuda":
        raise RuntimeError("canny_cuda needs a CUDA tensor (controlar_b200 has no CPU path)")
    if img.dim() == 2:
        img = img.unsqueeze(-1)
    x = img.to(torch.uint8).contiguous()
    H, W, Cc = x.shape
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        work = torch.empty(int(lib.car_canny_workspace_bytes(H, W)), dtype=torch.uint8, device=x.device)
        out = torch.empty(H, W, dtype=torch.uint8, device=x.device)
        changed = torch.zeros(1, dtype=torch.int32, device=x.device)
        restart = 1
        while True:
            check(lib.car_canny_u8(_ptr(x), H, W, Cc, int(math.floor(low_threshold)), int(math.floor(high_threshold)), _ptr(out), _ptr(work),
                                   sweeps_per_call, restart, _ptr(changed), cur_stream()), "car_canny_u8")
            restart = 0
            if int(changed.item()) == 0:
                return out


class CannyDetector:
    def __call__(self, img, low_threshold=100, high_threshold=200):
        """input: array or tensor (H,W,3); output: array (H,W) — reference condition/canny.py:7-14."""
        if not torch.is_tensor(img):
            img = torch.from_numpy(np.ascontiguousarray(img))
        dev = img.device if img.device.type == "cuda" else torch.device("cuda")
        return canny_cuda(img.to(dev), low_threshold, high_threshold).cpu().numpy()
