"""Soft-classification encode/decode facades (reference: src/spe/classification_utils.py).

Histogram construction and label *encoding* are one-off / dataset-side host work in the reference and stay
NumPy here (vectorised); *decoding* -- the hot path -- runs in libspef_b200.so (spef_decode_ori / spef_decode_pos).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .. import _ffi
from ..engine import Engine

_ORI_MIN = np.array([-180, -90, -180])
_ORI_MAX = np.array([180, 90, 180])


def _grid(n: int) -> np.ndarray:
    g = np.linspace(0.0, 1.0, n)
    return np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)


def _euler_to_quat(yaw, pitch, roll):
    """Vectorised scalar-first ZYX conversion, degrees (src/spe/utils.py:211-230)."""
    hy, hp, hr = np.deg2rad(yaw) / 2, np.deg2rad(pitch) / 2, np.deg2rad(roll) / 2
    cy, sy, cp, sp, cr, sr = np.cos(hy), np.sin(hy), np.cos(hp), np.sin(hp), np.cos(hr), np.sin(hr)
    q = np.stack([cy * cp * cr + sy * sp * sr, cy * cp * sr - sy * sp * cr,
                  cy * sp * cr + sy * cp * sr, sy * cp * cr - cy * sp * sr], axis=-1)
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


class _DeviceDecoder:
    """Lazily created post-processing context (no network weights needed for decode / score)."""

    def __init__(self):
        self._engine: Optional[Engine] = None

    def _post_engine(self, n_ori: int = 8, n_pos: int = 3) -> Engine:
        if self._engine is None:
            self._engine = Engine(32, 32, n_ori, n_pos, False, "fp32", 1)
        return self._engine


class OrientationSoftClassification(_DeviceDecoder):
    """classification_utils.py:10-176."""

    def __init__(self, n_bins_per_dim: int, smooth_factor: int, delete_unused_bins: bool = False):
        super().__init__()
        self.n_bins_per_dim = n_bins_per_dim
        self.smooth_factor = smooth_factor
        self.delete_unused_bins = delete_unused_bins
        self.histogram, self.redundant_flags = self.build_histogram(_ORI_MIN, _ORI_MAX)
        self.n_bins = self.histogram.shape[0]
        self._b = None

    def build_histogram(self, min_lim: np.ndarray, max_lim: np.ndarray):
        """:39-83 -- yaw slowest, roll fastest; redundant = yaw==180 or roll==180, or |pitch|==90 and yaw!=-180."""
        eul = _grid(self.n_bins_per_dim) * (max_lim - min_lim) + min_lim
        q = _euler_to_quat(eul[:, 0], eul[:, 1], eul[:, 2])
        boundary = np.logical_or(eul[:, 0] == max_lim[0], eul[:, 2] == max_lim[2])
        gimbal = np.logical_and(np.abs(eul[:, 1]) == max_lim[1], eul[:, 0] != min_lim[0])
        red = np.logical_or(boundary, gimbal)
        if self.delete_unused_bins:
            q = q[~red]
        return q, red

    @property
    def b(self) -> np.ndarray:
        """Pre-computed outer products q q^T [n,4,4] (:168-176); kept for API compatibility, the kernel forms them
        on the fly."""
        if self._b is None:
            self._b = self.histogram[:, :, None] * self.histogram[:, None, :]
        return self._b

    def encode(self, ori: np.ndarray) -> np.ndarray:
        """:85-111."""
        variance = (self.smooth_factor / self.n_bins_per_dim) ** 2 / 12
        c = np.minimum(1.0, np.abs(self.histogram @ np.asarray(ori, np.float64)))
        k = np.exp(-((2 * np.arccos(c) / np.pi) ** 2) / (2 * variance))
        if not self.delete_unused_bins:
            k[self.redundant_flags] = 0
        p = k / np.sum(k)
        if np.any(np.isnan(p)):
            raise ValueError('NaN found in encoded orientation')
        return p.astype(np.float32)

    def _engine_ready(self) -> Engine:
        eng = self._post_engine(self.n_bins)
        if eng.ori_hist_n != self.n_bins:
            eng.set_ori_histogram(self.histogram)
        return eng

    def encode_batch(self, ori_batch: np.ndarray) -> np.ndarray:
        """[B,4] true orientations -> [B,n_bins] float32 pdfs on the GPU (spef_encode_ori): `encode` (:85-111) for a batch,
        e.g. label generation for a whole dataset.  Raises like `encode` when a pdf has a NaN."""
        import torch
        variance = (self.smooth_factor / self.n_bins_per_dim) ** 2 / 12
        masked = None if self.delete_unused_bins else torch.from_numpy(self.redundant_flags.astype(np.uint8))
        out, flags = self._engine_ready().encode_ori(torch.as_tensor(np.asarray(ori_batch, np.float64)), variance, masked)
        if bool((flags & _ffi.FLAG_ENC_NAN).any()):
            raise ValueError('NaN found in encoded orientation')
        return out.cpu().numpy()

    def decode_batch(self, ori_batch: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """:149-166 -- [B,n_bins] pdfs -> ([B,4] float32 quaternions, [B,4,4] float32 inv(A))."""
        out = self._engine_ready().decode_ori_host(np.asarray(ori_batch), is_logits=False, want_hinv=True)
        if np.any(out["flags"] & _ffi.FLAG_ORI_NAN):
            raise ValueError("Error during orientation decoding")
        return out["quat"], out["hinv"]

    def decode(self, ori: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """:113-147."""
        q, h = self.decode_batch(np.asarray(ori).reshape(1, -1))
        return q[0], h[0]

    def pre_compute_ori_decode(self) -> np.ndarray:
        return self.b


class PositionSoftClassification(_DeviceDecoder):
    """classification_utils.py:179-285."""

    def __init__(self, n_bins_per_dim: int, smooth_factor: int, min_lim: np.ndarray, max_lim: np.ndarray):
        super().__init__()
        self.n_bins_per_dim = n_bins_per_dim
        self.smooth_factor = smooth_factor
        self.min_lim = min_lim
        self.max_lim = max_lim
        self.histogram = self.build_histogram()
        self.n_bins = self.histogram.shape[0]

    def build_histogram(self) -> np.ndarray:
        return _grid(self.n_bins_per_dim) * (self.max_lim - self.min_lim) + self.min_lim

    def encode(self, pos: np.ndarray) -> np.ndarray:
        """:218-240."""
        variance = (self.smooth_factor / self.n_bins_per_dim) ** 2 / 12
        k = np.exp(-np.sum((np.asarray(pos, np.float64) - self.histogram) ** 2, axis=1) / (2 * variance))
        p = k / np.sum(k)
        if np.any(np.isnan(p)):
            raise ValueError('NaN found in encoded position')
        return p.astype(np.float32)

    def _engine_ready(self) -> Engine:
        eng = self._post_engine(8, 3)
        if eng.pos_hist_n != self.n_bins:
            eng.set_pos_histogram(self.histogram)
        return eng

    def encode_batch(self, pos_batch: np.ndarray) -> np.ndarray:
        """[B,3] true positions -> [B,n_bins] float32 pdfs on the GPU (spef_encode_pos): `encode` (:218-240) for a batch."""
        import torch
        variance = (self.smooth_factor / self.n_bins_per_dim) ** 2 / 12
        out, flags = self._engine_ready().encode_pos(torch.as_tensor(np.asarray(pos_batch, np.float64)), variance)
        if bool((flags & _ffi.FLAG_ENC_NAN).any()):
            raise ValueError('NaN found in encoded position')
        return out.cpu().numpy()

    def decode_batch(self, pos_batch: np.ndarray) -> np.ndarray:
        """:269-285."""
        out = self._engine_ready().decode_pos_host(np.asarray(pos_batch), is_logits=False)
        if np.any(out["flags"] & _ffi.FLAG_POS_ZERO_SUM):
            raise ValueError("Encoded position vector sum is zero, cannot decode.")
        if np.any(out["flags"] & _ffi.FLAG_POS_NAN):
            raise ValueError("Error during position decoding, NaN found in decoded position.")
        return out["pos"]

    def decode(self, pos: np.ndarray) -> np.ndarray:
        """:242-267."""
        return self.decode_batch(np.asarray(pos).reshape(1, -1))[0]
