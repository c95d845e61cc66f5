"""SPEUtils facade (reference: src/spe/spe_utils.py:10-159): softmax / decode / score through libspef_b200.so."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import _ffi
from ..engine import Engine
from .classification_utils import OrientationSoftClassification, PositionSoftClassification

_score_engine: Optional[Engine] = None


def _get_score_engine() -> Engine:
    global _score_engine
    if _score_engine is None:
        _score_engine = Engine(32, 32, 8, 3, False, "fp32", 1)
    return _score_engine


class SPEUtils:
    """Spacecraft Pose Estimation Utils -- same constructor and attributes as the reference (:15-54)."""

    def __init__(self, camera, ori_mode: str = 'regression', n_ori_bins_per_dim: int = 12, ori_smooth_factor: int = 3,
                 ori_delete_unused_bins: bool = True, pos_mode: str = 'regression', n_pos_bins_per_dim: int = 10,
                 pos_smooth_factor: int = 100, keypoints_path: str = None):
        assert ori_mode in ['regression', 'classification', 'keypoints']
        assert pos_mode in ['regression', 'classification', 'keypoints']
        if pos_mode == 'keypoints' or ori_mode == 'keypoints' or keypoints_path is not None:
            raise NotImplementedError("keypoint mode is outside the B200 hot path (SURVEY.md section 2, row 9)")
        self.ori_mode = ori_mode
        self.pos_mode = pos_mode
        self.camera = camera
        self.orientation = OrientationSoftClassification(n_ori_bins_per_dim, ori_smooth_factor, ori_delete_unused_bins)
        self.position = PositionSoftClassification(n_pos_bins_per_dim, pos_smooth_factor,
                                                   min_lim=np.array([-16, -12, -2]), max_lim=np.array([16, 12, 40]))
        self.keypoints = None

    def last_activ(self, pose: dict) -> dict:
        """:56-81 -- softmax over the classification logits (float32, max-subtracted)."""
        if self.ori_mode == 'regression':
            raise NotImplementedError("orientation regression is outside the B200 hot path (soft-classification only)")
        out = self.orientation._engine_ready().decode_ori_host(pose['ori_soft'], is_logits=True, want_soft=True)
        pose['ori_soft'] = out["soft"]
        if self.pos_mode == 'classification':
            out = self.position._engine_ready().decode_pos_host(pose['pos_soft'], is_logits=True, want_soft=True)
            pose['pos_soft'] = out["soft"]
        return pose

    def decode(self, pose: dict) -> dict:
        """:83-101."""
        if self.ori_mode == 'classification':
            pose['ori'], _ = self.orientation.decode_batch(pose['ori_soft'])
        if self.pos_mode == 'classification':
            pose['pos'] = self.position.decode_batch(pose['pos_soft'])
        return pose

    @staticmethod
    def get_score(true_pose: dict, pred_pose: dict) -> dict:
        """:104-159 -- ESA score of a batch.  The `> 1.01` ValueError of :137-138 is dead code in the reference
        (it can never fire), so -- like the reference -- values above 1 are clamped and nothing is raised."""
        sums, _ = _get_score_engine().score_host(pred_pose['ori'], pred_pose['pos'], true_pose['ori'], true_pose['pos'])
        return SPEUtils.metrics_from_sums(sums)

    @staticmethod
    def metrics_from_sums(sums: np.ndarray) -> dict:
        """Batch means from the kernel's float64 sums, in the reference's float32 scalar arithmetic (:126-149)."""
        n = sums[3]
        ori = np.float32(sums[0] / n)
        pos = np.float32(sums[1] / n)
        return {
            'esa_score': ori + pos,
            'ori_score': ori,
            'pos_score': pos,
            'ori_error': ori * 180 / np.pi,
            'pos_error': np.float32(sums[2] / n),
        }
