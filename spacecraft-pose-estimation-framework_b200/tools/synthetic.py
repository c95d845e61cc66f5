"""Synthetic, seed-deterministic inputs for benchmarks and parity tests (there is no network for datasets or
checkpoints): SPEED-shaped images, pose targets and the "calibrated random init" of SURVEY.md section 8(d).

Why not the reference's default init: kaiming-normal(fan_out) on depthwise weights with BatchNorm running stats
(0, 1) shrinks activations ~10x per stage; the logits come out ~1e-10 and the softmax is exactly uniform, which
makes any parity test vacuous.  Here conv weights are He-normal (fan_in), BN affine parameters are random, and the
BN running statistics are the ones measured once by a train-mode pass over torch.rand(8,3,240,384)
(tests/golden/make_goldens.py --calibrate) and committed as data/bn_calib_seed7.npz, so that every machine builds
bit-identical weights without running the network.  Only RNG draws and file reads happen here -- no inference.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Tuple

import numpy as np
import torch

from ..modeling import arch

_DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
BN_CALIB_PATH = os.path.join(_DATA_DIR, "bn_calib_seed7.npz")
WEIGHT_SEED = 7
IMAGE_SEED = 1001  # eval.py:14 of the reference


def raw_state_dict(n_ori: int = 1728, n_pos: int = 3, seed: int = WEIGHT_SEED) -> Dict[str, torch.Tensor]:
    """Random weights and BN affine parameters; BN running stats left at (0, 1).  Deterministic in `seed`."""
    g = torch.Generator().manual_seed(seed)
    residual_projects = {l["prefix"] for l in arch.conv_layers() if l["residual"]}
    sd = {}
    for key, shape, role in arch.state_dict_spec(n_ori, n_pos):
        prefix = key.rsplit(".", 2)[0]
        if role == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif role == "bn_weight":
            lo, hi = (0.15, 0.45) if prefix in residual_projects else (0.5, 1.5)  # tames the net's chaos
            sd[key] = torch.rand(shape, generator=g) * (hi - lo) + lo
        elif role == "bn_bias":
            sd[key] = torch.randn(shape, generator=g) * 0.1
        elif role == "bn_mean":
            sd[key] = torch.zeros(shape)
        elif role == "bn_var":
            sd[key] = torch.ones(shape)
        elif role == "bn_count":
            sd[key] = torch.tensor(1, dtype=torch.long)
        elif role == "linear_weight":
            sd[key] = torch.randn(shape, generator=g) * 0.3  # N(0, 0.01) * 30: logits std ~5
        elif role == "linear_bias":
            sd[key] = torch.zeros(shape)
    if n_pos == 3:
        sd["head.pos.0.bias"] = torch.tensor([0.0, 0.0, 10.0])
    return sd


def synthetic_state_dict(n_ori: int = 1728, n_pos: int = 3, seed: int = WEIGHT_SEED) -> Dict[str, torch.Tensor]:
    """The calibrated random init: raw_state_dict + committed BN running statistics."""
    if seed != WEIGHT_SEED:
        raise ValueError(f"BN calibration is committed for seed {WEIGHT_SEED} only")
    if not os.path.isfile(BN_CALIB_PATH):
        raise FileNotFoundError(f"{BN_CALIB_PATH} missing: run `python tests/golden/make_goldens.py --calibrate`")
    sd = raw_state_dict(n_ori, n_pos, seed)
    calib = np.load(BN_CALIB_PATH)
    for l in arch.conv_layers():
        p = l["prefix"]
        sd[p + ".1.running_mean"] = torch.from_numpy(calib[p + ".mean"].astype(np.float32))
        sd[p + ".1.running_var"] = torch.from_numpy(calib[p + ".var"].astype(np.float32))
    return sd


def synthetic_images(batch: int, img_size: Tuple[int, int] = (240, 384), seed: int = IMAGE_SEED) -> torch.Tensor:
    """[B,3,H,W] float32 in [0,1): the reference tensor contract (Resize -> ToTensor, no mean/std normalisation,
    src/data/datasets/speed.py:66-69).  White noise keeps the calibrated net well conditioned."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((batch, 3, img_size[0], img_size[1]), generator=g)


def synthetic_targets(batch: int, seed: int = 2024) -> Dict[str, np.ndarray]:
    """Uniform random unit quaternions (Shoemake, as src/spe/utils.py:415-447) and SPEED-like positions
    z ~ U(3, 35) m, x, y ~ U(-0.3 z, 0.3 z) (create_dspeed.py:69-82)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand((batch, 6), generator=g).double().numpy()
    x0, t1, t2 = u[:, 0], 2 * np.pi * u[:, 1], 2 * np.pi * u[:, 2]
    r1, r2 = np.sqrt(1 - x0), np.sqrt(x0)
    q = np.stack([np.sin(t1) * r1, np.cos(t1) * r1, np.sin(t2) * r2, np.cos(t2) * r2], axis=1)
    z = 3 + 32 * u[:, 3]
    pos = np.stack([(2 * u[:, 4] - 1) * 0.3 * z, (2 * u[:, 5] - 1) * 0.3 * z, z], axis=1)
    return {"ori": q.astype(np.float32), "pos": pos.astype(np.float32)}


class SyntheticLoader:
    """List-of-batches stand-in for the reference DataLoader: yields ({'torch': images}, {'ori','pos'}) like
    SPEDataset (src/data/utils.py:212-249).  `rank`/`world` shard the batches round-robin for multi-GPU evaluation."""

    def __init__(self, n_images: int, batch_size: int, img_size=(240, 384), seed: int = IMAGE_SEED, rank: int = 0,
                 world: int = 1, pin: bool = False, reuse_images: bool = False):
        self.batches = []
        tg = synthetic_targets(n_images, seed + 17)
        base = synthetic_images(batch_size, img_size, seed) if reuse_images else None
        for bi, start in enumerate(range(0, n_images, batch_size)):
            if bi % world != rank:
                continue
            n = min(batch_size, n_images - start)
            img = base[:n] if reuse_images else synthetic_images(n, img_size, seed + 1 + bi)
            if pin and torch.cuda.is_available():
                img = img.pin_memory()
            tgt = {"ori": torch.from_numpy(tg["ori"][start:start + n]), "pos": torch.from_numpy(tg["pos"][start:start + n])}
            self.batches.append(({"torch": img}, tgt))

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def synthetic_frames(batch: int, height: int = 1200, width: int = 1920, channels: int = 1, seed: int = 4242,
                     kind: str = "speed") -> np.ndarray:
    """uint8 camera frames [B,H,W] (channels = 1, SPEED's greyscale JPEGs) or [B,H,W,3], the input of
    SPEDataset.__getitem__ before Resize -> ToTensor (src/data/utils.py:212-226; Camera 1920 x 1200,
    src/data/datasets/speed.py:23-24).  kind 'speed': dark sky, sensor noise and one bright textured blob per frame;
    'noise': uniform white noise (worst case for the resampling filter)."""
    rng = np.random.default_rng(seed)
    shape = (batch, height, width) if channels == 1 else (batch, height, width, channels)
    if kind == "noise":
        return rng.integers(0, 256, size=shape, dtype=np.uint8)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    out = np.empty(shape, np.uint8)
    for b in range(batch):
        cy, cx = rng.uniform(0.2, 0.8) * height, rng.uniform(0.2, 0.8) * width
        s = rng.uniform(0.03, 0.15) * min(height, width)
        blob = 230.0 * np.exp(-(((yy - cy) / s) ** 2 + ((xx - cx) / (1.4 * s)) ** 2))
        tex = 0.75 + 0.25 * np.sin(xx * 0.9 + b) * np.cos(yy * 0.7)
        for c in range(channels):
            f = blob * tex * (1.0 - 0.1 * c) + rng.normal(6.0, 3.0, size=(height, width))
            plane = np.clip(np.rint(f), 0, 255).astype(np.uint8)
            if channels == 1:
                out[b] = plane
            else:
                out[b, :, :, c] = plane
    return out
