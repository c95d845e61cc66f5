"""Inference (reference: src/temporal/inference.py:20-195): per-frame pose + adaptive temporal pdf filtering.

The whole frame step -- forward, softmax, still decode, sign continuity, the two TemporalPDF filters, filtered
decode, sign continuity -- is one C-ABI call (spef_temporal_step); filter and sign state live in the device
context.  `n_streams` > 1 runs that many independent videos in lock-step (one frame of each per call)."""
from __future__ import annotations

import time
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _ffi
from ..modeling.model import MobileURSONetB200
from ..spe.spe_utils import SPEUtils


class Inference:
    def __init__(self, model: MobileURSONetB200, inference_device: str, spe_utils: SPEUtils, n_streams: int = 1):
        self.model = model
        self.inference_device = inference_device
        self.spe_utils = spe_utils
        self.n_streams = n_streams
        self.img_size = None
        self.engine = None
        self.select_inference_engine(inference_device)

    def select_inference_engine(self, device: str, model_name: str = None):
        """inference.py:46-80.  Only the host GPU is a B200 back-end; the reference's other targets (host CPU,
        Jetson, Ultra96) are different deployments and out of scope."""
        assert device in ('gpu_host', 'cpu_host', 'gpu_jetson', 'cpu_ultra96')
        if device != 'gpu_host':
            raise NotImplementedError(f"inference device '{device}' is not a B200 back-end (use the reference for it)")
        assert torch.cuda.is_available()
        self.inference_device = device
        dev = torch.device('cuda', torch.cuda.current_device())
        self.model.to(dev)
        self.model.eval()
        eng = self.model.engine(dev)
        eng.set_ori_histogram(self.spe_utils.orientation.histogram)
        if self.model.pos_classification:
            eng.set_pos_histogram(self.spe_utils.position.histogram)
        self.engine = eng
        self._need_reset = True

    def reset(self) -> None:
        """inference.py:92-99."""
        self._need_reset = True

    def update(self, model, spe_utils) -> None:
        """inference.py:101-112."""
        self.model = model
        self.spe_utils = spe_utils
        self.select_inference_engine(self.inference_device)
        self.reset()

    def predict(self, image: torch.Tensor, video_type: str = None) -> Tuple[dict, float, Optional[dict]]:
        """inference.py:114-195.  image: [n_streams,3,H,W].  Returns (pose_still, latency_ms, pose_video | None);
        with n_streams == 1 the dict values have no batch dimension, exactly like the reference."""
        if not self.img_size or self.img_size != tuple(image.size()):
            self.img_size = tuple(image.size())
        if video_type is not None and video_type != 'Adaptative':
            raise ValueError(f'type of video filtering not implemented: {video_type}')
        if video_type == 'Adaptative':
            assert self.spe_utils.ori_mode == 'classification'
            assert self.spe_utils.pos_mode == 'classification'
        if not self.model.pos_classification:
            raise NotImplementedError("Inference on the device needs the Mobile-URSONet+ heads (classification position)")
        S = image.shape[0]
        if self._need_reset or getattr(self, "_streams", None) != S:
            self.engine.temporal_reset(S)
            self._streams = S
            self._need_reset = False
        t1 = time.time()
        out = self.engine.temporal_step(image, apply_filter=(video_type is not None))
        out = {k: v.cpu().numpy() for k, v in out.items()}
        t2 = time.time()
        flags = out["flags"]
        if np.any(flags & _ffi.FLAG_ORI_NAN):
            raise ValueError("Error during orientation decoding")
        if np.any(flags & _ffi.FLAG_POS_ZERO_SUM):
            raise ValueError("Encoded position vector sum is zero, cannot decode.")
        if np.any(flags & _ffi.FLAG_POS_NAN):
            raise ValueError("Error during position decoding, NaN found in decoded position.")

        def sq(a):
            return a[0] if S == 1 else a

        pose_still = {'ori_soft': sq(out['still_ori_soft']), 'pos_soft': sq(out['still_pos_soft']),
                      'ori': sq(out['still_quat']), 'pos': sq(out['still_pos'])}
        pose_video = None
        if video_type is not None:
            pose_video = {'ori_soft': sq(out['video_ori_soft']), 'ori_distance': sq(out['ori_distance']),
                          'pos_soft': sq(out['video_pos_soft']), 'pos_distance': sq(out['pos_distance']),
                          'ori': sq(out['video_quat']), 'pos': sq(out['video_pos'])}
        return pose_still, (t2 - t1) * 1000, pose_video
