"""Model construction API of the reference (src/modeling/model.py:70-279) backed by the B200 engine.

``import_model`` keeps the reference's signature and return value ``(model, bit_width)``.  The returned
``MobileURSONetB200`` is an ``nn.Module`` only as a *parameter container* with the reference's 316
state_dict keys (``features.features.{i}...``, ``head.ori.1.*``, ``head.pos.0.*``); its ``forward`` is one
call into libspef_b200.so (spef_forward) and it owns no torch compute modules, so there is no eager fallback.
It can be handed to the reference's own ``SPETorch`` (any callable ``model(x) -> (ori, pos)`` with
``.to()``/``.eval()``, src/spe/spe_torch.py:24-39) or to ``spef_b200.spe.SPEB200`` (fused predict).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import arch
from ..engine import Engine

_SUPPORTED_BACKBONES = ("mobilenet_v2_pytorch",)
_SUPPORTED_HEADS = ("ursonet_pytorch",)


class _Container(nn.Module):
    """Name-space node of the parameter tree (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the network runs inside libspef_b200.so, not in torch")


class MobileURSONetB200(nn.Module):
    """Mobile-URSONet (MobileNetV2 + URSONet head) whose forward runs on hand-written sm_100a kernels."""

    def __init__(self, n_ori_outputs: int, n_pos_outputs: int, pos_classification: bool = False,
                 img_size=(240, 384), precision: str = "bf16", max_batch: int = 32, pw_impl: int = 0):
        super().__init__()
        self.n_ori, self.n_pos = int(n_ori_outputs), int(n_pos_outputs)
        self.pos_classification = bool(pos_classification)
        self.img_size = (int(img_size[0]), int(img_size[1]))
        self.precision, self.max_batch, self.pw_impl = precision, int(max_batch), int(pw_impl)
        self.features = _Container()
        self.head = _Container()
        self._register_tree()
        self.reset_parameters()
        self._engine: Optional[Engine] = None
        self._weights_version = 0      # bumped whenever parameters may have changed
        self._engine_version = -1

    # ---- parameter tree with the reference's keys ----------------------------------------------------
    def _register_tree(self):
        for key, shape, role in arch.state_dict_spec(self.n_ori, self.n_pos):
            parts = key.split(".")
            node = self
            for p in parts[:-1]:
                if p not in node._modules:
                    node.add_module(p, _Container())
                node = node._modules[p]
            if role in ("bn_mean", "bn_var"):
                node.register_buffer(parts[-1], torch.zeros(shape) if role == "bn_mean" else torch.ones(shape))
            elif role == "bn_count":
                node.register_buffer(parts[-1], torch.tensor(0, dtype=torch.long))
            else:
                node.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))

    def reset_parameters(self):
        """ModelWrapper's initialisation (src/modeling/common/pytorch_layers.py:17-27): conv kaiming-normal
        (fan_out), BN weight 1 / bias 0, linear N(0, 0.01) / bias 0."""
        sd = self.state_dict()
        with torch.no_grad():
            for key, shape, role in arch.state_dict_spec(self.n_ori, self.n_pos):
                t = sd[key]
                if role == "conv":
                    fan_out = shape[0] * shape[2] * shape[3]
                    t.normal_(0.0, math.sqrt(2.0 / fan_out))
                elif role == "bn_weight":
                    t.fill_(1.0)
                elif role in ("bn_bias", "linear_bias", "bn_mean"):
                    t.zero_()
                elif role == "bn_var":
                    t.fill_(1.0)
                elif role == "linear_weight":
                    t.normal_(0.0, 0.01)
        self._weights_version = getattr(self, "_weights_version", 0) + 1

    # ---- nn.Module plumbing --------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._weights_version += 1
        return res

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .cpu() / .float(): parameters stay f32; note a possible device move
        out = super()._apply(fn, *a, **k)
        self._weights_version += 1
        return out

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("spef_b200 implements the inference path only (BatchNorm is folded); "
                                      "training is out of scope (SURVEY.md section 2, row 13)")
        return super().train(False)

    # ---- engine --------------------------------------------------------------------------------------
    def engine(self, device=None) -> Engine:
        """The device context for this model; (re)built when the device changes, weights re-uploaded when they
        changed."""
        if device is None or torch.device(device).type != "cuda":
            p = next(self.parameters())
            device = p.device if p.is_cuda else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        device = torch.device(device)
        index = device.index if device.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
        device = torch.device("cuda", index)
        if self._engine is not None and self._engine.device != device:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine = Engine(self.img_size[0], self.img_size[1], self.n_ori, self.n_pos, self.pos_classification,
                                  self.precision, self.max_batch, device, self.pw_impl)
            self._engine_version = -1
        if self._engine_version != self._weights_version:
            self._engine.load_state_dict(self.state_dict())
            self._engine_version = self._weights_version
        return self._engine

    def release_engine(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

    def forward(self, x: torch.Tensor):
        """ModelWrapper.forward (pytorch_layers.py:29-32): x [B,3,H,W] float32 in [0,1] -> (ori, pos) on x's device
        when x is a CUDA tensor, else on the engine's device (SPETorch moves inputs to the model's device itself)."""
        eng = self.engine(x.device if x.device.type == "cuda" else None)
        return eng.forward(x)


def save_model(model: nn.Module, bit_width: Optional[dict], path: str) -> None:
    """src/modeling/model.py:70-89: parameters.pt (+ bit_width.json for Brevitas models, never the case here)."""
    os.makedirs(path, exist_ok=True)
    torch.save(model.state_dict(), os.path.join(path, "parameters.pt"))


def copy_state_dict(state_dict_1: Dict[str, torch.Tensor], state_dict_2: Dict[str, torch.Tensor], act_quant: bool = False) -> Dict[str, torch.Tensor]:
    """Manual copy of state_dict_1 into state_dict_2 for state dicts whose keys differ (src/modeling/model.py:92-119): per key
    FAMILY ("weight", "bias", "running_mean", "running_var", "num_batches_tracked", and "act_quant" when asked), the i-th tensor
    of that family in state_dict_1 goes to the i-th key of that family in state_dict_2; other keys (e.g. the act_quant entries
    of a Brevitas checkpoint when act_quant=False) are ignored, as in the reference.  One addition: a shape check, because a
    silent mis-ordered copy would only surface as wrong poses."""
    keys1, keys2 = list(state_dict_1.keys()), list(state_dict_2.keys())
    families = ["weight", "bias", "running_mean", "running_var", "num_batches_tracked"]
    if act_quant:
        families.append("act_quant")
    for fam in families:
        k1 = [k for k in keys1 if fam in k]
        k2 = [k for k in keys2 if fam in k]
        if len(k1) > len(k2):
            raise ValueError(f"copy_state_dict: {len(k1)} '{fam}' tensors in the source but only {len(k2)} in the destination")
        for i, src_key in enumerate(k1):
            if tuple(state_dict_1[src_key].shape) != tuple(state_dict_2[k2[i]].shape):
                raise ValueError(f"shape mismatch copying {src_key} -> {k2[i]}: {tuple(state_dict_1[src_key].shape)} vs {tuple(state_dict_2[k2[i]].shape)}")
            state_dict_2[k2[i]] = state_dict_1[src_key]
    return state_dict_2


def import_model(
    data: dict,
    backbone_name: str,
    head_name: str,
    params_path: str = None,
    bit_width_path: str = None,
    manual_copy: bool = False,
    in_channels: int = 3,
    batchnorm: bool = True,
    residual: bool = True,
    quantization: bool = True,
    ori_mode: str = 'classification',
    n_ori_bins: int = None,
    pos_mode: str = 'regression',
    n_pos_bins: int = None,
    *,
    precision: str = "fp32",
    max_batch: int = None,
    pw_impl: int = 0,
) -> tuple:
    """Same contract as the reference's import_model (src/modeling/model.py:122-279); returns (model, bit_width)
    with bit_width = None for PyTorch backbones (:180-182).  Keyword-only extras select the B200 engine's
    precision and workspace size.  precision defaults to 'fp32' -- the reference model family is FP32 and a drop-in caller gets
    its arithmetic (logits within 1e-4 relative) unless it asks for the BF16 tensor-core path with precision='bf16'
    (the throughput path of bench.py).

    Differences, all outside the inference arithmetic: only the FP32 PyTorch model family is built
    ('mobilenet_v2_pytorch' + 'ursonet_pytorch'; Brevitas/FINN variants are out of scope), the dry-run forward
    (:259) is skipped (it exists for Brevitas quantiser state), and without `params_path` the reference's random
    initialisation is applied (the ImageNet download at :268-277 cannot work offline and is swallowed there too).
    """
    assert ori_mode in ['classification', 'regression', 'keypoints']
    assert pos_mode in ['classification', 'regression', 'keypoints']
    if ori_mode == 'classification':
        assert n_ori_bins is not None
    if pos_mode == 'classification':
        assert n_pos_bins is not None
    if backbone_name not in _SUPPORTED_BACKBONES:
        raise NotImplementedError(f"backbone '{backbone_name}' is outside the B200 hot path (supported: {_SUPPORTED_BACKBONES})")
    if head_name not in _SUPPORTED_HEADS:
        raise NotImplementedError(f"head '{head_name}' is outside the B200 hot path (supported: {_SUPPORTED_HEADS})")
    if ori_mode != 'classification':
        raise NotImplementedError("the B200 path implements the soft-classification orientation head (ori_mode='classification')")
    if pos_mode == 'keypoints':
        raise NotImplementedError("keypoint mode is out of scope (SURVEY.md section 2, row 9)")
    if in_channels != 3 or not batchnorm or not residual:
        raise NotImplementedError("the B200 kernels implement the default topology: in_channels=3, batchnorm=True, residual=True")

    # Example batch: only the image size is needed (model.py:170)
    image, _target = next(iter(data[list(data.keys())[0]]))
    img = image['torch']
    img_size = (int(img.size(2)), int(img.size(3)))
    batch_hint = int(img.size(0))

    model = MobileURSONetB200(
        n_ori_outputs=n_ori_bins,
        n_pos_outputs=3 if pos_mode == 'regression' else n_pos_bins,
        pos_classification=(pos_mode == 'classification'),
        img_size=img_size, precision=precision,
        max_batch=max_batch if max_batch is not None else max(batch_hint, 32), pw_impl=pw_impl)

    if params_path is not None:
        assert os.path.isfile(params_path), f'Parameters not found {params_path}'
        sd = torch.load(params_path, map_location="cpu")
        if manual_copy:
            model.load_state_dict(copy_state_dict(sd, model.state_dict()))
        else:
            model.load_state_dict(sd)
    model.eval()
    return model, None
