"""SPEB200 -- the predict() plug-in (reference duck type: SPETorch, src/spe/spe_torch.py:12-124)."""
from __future__ import annotations

import gc
import time
from typing import Dict, Tuple

import numpy as np
import torch

from .. import _ffi
from ..modeling.model import MobileURSONetB200
from .spe_utils import SPEUtils


class SPEB200:
    """Drop-in for SPETorch: predict(images) -> (pose dict of float32 NumPy arrays, latency_ms).

    forward + softmax + decode run as one stream-ordered sequence of sm_100a kernels (spef_predict_host for CPU
    tensors, spef_predict for CUDA tensors)."""

    def __init__(self, model: MobileURSONetB200, device: torch.device, spe_utils: SPEUtils, lanes: int = 2) -> None:
        self.model = model
        self.device = torch.device(device)
        self.spe_utils = spe_utils
        # evaluation() issues the batches of a phase round-robin over this many device contexts / CUDA streams (Engine.lanes:
        # independent batches overlap on the GPU, +7 % images/s at batch 256); predict() itself is one call on one context
        self.lanes = max(1, int(lanes))
        self._bind()

    def _bind(self):
        if not isinstance(self.model, MobileURSONetB200):
            raise TypeError("SPEB200 drives a spef_b200 model (import_model); wrap other models in the reference's SPETorch")
        if self.device.type != "cuda":
            raise RuntimeError("SPEB200 runs on a CUDA device (B200); there is no CPU path")
        if self.spe_utils.ori_mode != 'classification':
            raise NotImplementedError("SPEB200 implements ori_mode='classification'")
        if (self.spe_utils.pos_mode == 'classification') != self.model.pos_classification:
            raise ValueError("spe_utils.pos_mode does not match the model's position head")
        self.model.to(self.device)
        self.model.eval()
        eng = self.model.engine(self.device)
        if self.spe_utils.orientation.n_bins != self.model.n_ori:
            raise ValueError(f"orientation histogram has {self.spe_utils.orientation.n_bins} bins, model head has {self.model.n_ori}")
        eng.set_ori_histogram(self.spe_utils.orientation.histogram)
        if self.model.pos_classification:
            if self.spe_utils.position.n_bins != self.model.n_pos:
                raise ValueError(f"position histogram has {self.spe_utils.position.n_bins} bins, model head has {self.model.n_pos}")
            eng.set_pos_histogram(self.spe_utils.position.histogram)
        self.engine = eng

    def predict(self, images: torch.Tensor) -> Tuple[Dict, float]:
        """spe_torch.py:41-76.  Keys: 'ori_soft' [B,n_ori], 'ori' [B,4], 'pos' [B,3] (+ 'pos_soft' for a classification
        position head).  Raises ValueError on the reference's decode guards (classification_utils.py:134-135,253-254,262-263)."""
        assert hasattr(self, 'model') and self.model is not None
        t1 = time.time()
        if images.device.type == "cuda":
            out = {k: v.cpu().numpy() for k, v in self.engine.predict(images, want_soft=True).items()}
        else:
            out = self.engine.predict_host(images, want_soft=True)
        t2 = time.time()
        flags = out.pop("flags")
        if np.any(flags & _ffi.FLAG_ORI_NAN):
            raise ValueError("Error during orientation decoding")
        if np.any(flags & _ffi.FLAG_POS_ZERO_SUM):
            raise ValueError("Encoded position vector sum is zero, cannot decode.")
        if np.any(flags & _ffi.FLAG_POS_NAN):
            raise ValueError("Error during position decoding, NaN found in decoded position.")
        pose = {'ori_soft': out['ori_soft'], 'pos': out['pos'], 'ori': out['ori']}
        if 'pos_soft' in out:
            pose['pos_soft'] = out['pos_soft']
        return pose, (t2 - t1) * 1000

    def update_model(self, model, device: torch.device) -> None:
        """spe_torch.py:78-98."""
        self.delete_model()
        self.model = model
        self.device = torch.device(device)
        self._bind()

    def delete_model(self) -> None:
        """spe_torch.py:100-110."""
        if getattr(self, "model", None) is not None:
            self.model.release_engine()
            self.model.to(torch.device("cpu"))
        self.model = None
        self.engine = None
        gc.collect()
        torch.cuda.empty_cache()

    def move_to_cpu(self) -> None:
        """spe_torch.py:112-118: releases the device context; parameters go back to host memory."""
        self.model.release_engine()
        self.model.to(torch.device("cpu"))
        self.engine = None
        gc.collect()
        torch.cuda.empty_cache()

    def move_to_gpu(self) -> None:
        """spe_torch.py:120-124."""
        self._bind()
