"""Input side of the hot path: the transform of the reference's SPEDataset (src/data/utils.py:212-249) on the GPU.

  FrameTransform  <- transforms.Compose([transforms.Resize(img_size), transforms.ToTensor()]) (src/data/datasets/speed.py:59-62)
                     applied to Image.open(path).convert("RGB") (src/data/utils.py:215-226), for a batch of decoded frames

Decoding the JPEG stays with the caller (np.array(Image.open(path))); this directory also holds the package data
(bn_calib_seed7.npz, see tools/synthetic.py)."""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import torch

from ..engine import Engine


class FrameTransform:
    """frames (uint8 [B,H,W] greyscale or [B,H,W,3] RGB, NumPy or torch, host or device) -> [B,3,h,w] on the engine's device:
    float32 in [0,1] exactly as the reference's loader builds it (default), or the uint8 pixels ToTensor divides by 255
    (dtype=torch.uint8; feed them to an engine set to uint8 images -- same results, 4x fewer bytes).  Bit-exact against
    torchvision + Pillow (libspef_b200: spef_resize_frames); there is no CPU fallback."""

    def __init__(self, engine: Engine, img_size: Tuple[int, int], dtype: torch.dtype = torch.float32):
        if tuple(img_size) != (engine.img_h, engine.img_w):
            raise ValueError(f"img_size {tuple(img_size)} does not match the engine's input size {(engine.img_h, engine.img_w)}")
        if dtype not in (torch.float32, torch.uint8):
            raise ValueError("dtype must be torch.float32 or torch.uint8")
        self.engine, self.dtype = engine, dtype

    def __call__(self, frames: Union[np.ndarray, torch.Tensor]) -> torch.Tensor:
        t = torch.from_numpy(frames) if isinstance(frames, np.ndarray) else frames
        if t.dim() == 2:   # one greyscale frame
            t = t.unsqueeze(0)
        return self.engine.resize_frames(t, self.dtype)
