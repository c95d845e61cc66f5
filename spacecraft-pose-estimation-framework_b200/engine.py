"""Thin Python owner of one ``spef_ctx`` (include/spef_b200.h).  PyTorch is used only for device memory
and stream handles; every computation is a call into libspef_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _ffi
from ._ffi import SpefConfig, SpefTemporalOut, check, ptr

_PRECISIONS = {"fp32": _ffi.SPEF_FP32, "float32": _ffi.SPEF_FP32, "bf16": _ffi.SPEF_BF16, "bfloat16": _ffi.SPEF_BF16}


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("spef_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    # no device given: the CURRENT device (one process per GPU sets it once; "cuda:0" would put every rank's helper contexts on GPU 0)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError(f"spef_b200 runs on CUDA devices only, got {dev}")
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


def _stream(dev: torch.device) -> Optional[int]:
    return torch.cuda.current_stream(dev).cuda_stream or None


class Engine:
    """One device context: weights, histograms, workspaces, temporal state."""

    def __init__(self, img_h: int = 240, img_w: int = 384, n_ori: int = 1728, n_pos: int = 3,
                 pos_classification: bool = False, precision: str = "bf16", max_batch: int = 32,
                 device=None, pw_impl: int = 0):
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.device = _require_cuda(device)
        self.lib = _ffi.lib()
        self.cfg = SpefConfig(C.sizeof(SpefConfig), self.device.index, img_h, img_w, n_ori, n_pos,
                              int(bool(pos_classification)), _PRECISIONS[precision], max_batch, pw_impl)
        self.precision = "bf16" if _PRECISIONS[precision] == _ffi.SPEF_BF16 else "fp32"
        self.act_dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        handle = C.c_void_p()
        check(None, self.lib.spef_create(C.byref(handle), C.byref(self.cfg)))
        self._h = handle.value
        self.n_ori, self.n_pos, self.pos_classification = n_ori, n_pos, bool(pos_classification)
        self.max_batch, self.img_h, self.img_w = max_batch, img_h, img_w
        self.weights_ready = False
        self.image_dtype = torch.float32
        self.ori_hist_n = self.pos_hist_n = 0
        # what a second lane (lanes()) needs to rebuild this context: constructor arguments, weights, tables
        self._ctor = (img_h, img_w, n_ori, n_pos, pos_classification, precision, max_batch, self.device, pw_impl)
        self._sd, self._sd_version = None, 0
        self._ori_hist, self._pos_hist, self._hist_version = None, None, 0
        self._twins: List["Engine"] = []
        self._twin_of = (-1, -1)          # (weights version, table version) of the parent this lane was synchronised to
        self.side_stream = None           # the CUDA stream lanes() gives this lane

    # ---- lifecycle -------------------------------------------------------------------------------
    def close(self):
        for t in getattr(self, "_twins", []):
            t.close()
        self._twins = []
        if getattr(self, "_h", None):
            self.lib.spef_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        check(self._h, rc)

    # ---- weights / tables ------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]):
        """Feed the reference's 316-key state_dict (src/modeling/model.py:261-266); BN folding, repacking and
        casting happen inside the library."""
        for key, t in state_dict.items():
            if key.endswith("num_batches_tracked"):
                continue
            a = t.detach().to("cpu", torch.float32).contiguous().numpy()
            shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
            self._ck(self.lib.spef_load_tensor(self._h, key.encode(), a.ctypes.data, shape, a.ndim))
        self._ck(self.lib.spef_finalize_weights(self._h))
        self.weights_ready = True
        self._sd, self._sd_version = state_dict, self._sd_version + 1

    def set_ori_histogram(self, hist: np.ndarray):
        h = np.ascontiguousarray(hist, dtype=np.float64)
        assert h.ndim == 2 and h.shape[1] == 4
        self._ck(self.lib.spef_set_ori_histogram(self._h, h.ctypes.data, h.shape[0]))
        self.ori_hist_n = h.shape[0]
        self._ori_hist, self._hist_version = h, self._hist_version + 1

    def set_pos_histogram(self, hist: np.ndarray):
        h = np.ascontiguousarray(hist, dtype=np.float64)
        assert h.ndim == 2 and h.shape[1] == 3
        self._ck(self.lib.spef_set_pos_histogram(self._h, h.ctypes.data, h.shape[0]))
        self.pos_hist_n = h.shape[0]
        self._pos_hist, self._hist_version = h, self._hist_version + 1

    def lanes(self, n: int = 2) -> List["Engine"]:
        """n device contexts with this one's weights, tables and image dtype, each with a CUDA stream of its own (side_stream):
        [self, twin, ...].  Independent batches issued round-robin over the lanes overlap on the GPU -- the last, partly filled
        wave of one batch's kernels (a persistent kernel ends when its slowest CTA does) runs next to the other batch's
        kernels instead of next to idle SMs, and launch gaps disappear: 138 k -> 148 k images/s at batch 256 with two lanes
        (a third adds nothing).  The twins are created on first use and re-synchronised when weights or tables changed; the
        ESA accumulators are per context (add the eval_read() vectors: the sums are linear)."""
        n = max(1, int(n))
        while len(self._twins) < n - 1:
            self._twins.append(Engine(*self._ctor))
        with torch.cuda.device(self.device):
            if self.side_stream is None:
                self.side_stream = torch.cuda.Stream(self.device)
            for t in self._twins[:n - 1]:
                if t.side_stream is None:
                    t.side_stream = torch.cuda.Stream(self.device)
                if t._twin_of[0] != self._sd_version and self._sd is not None:
                    t.load_state_dict(self._sd)
                if t._twin_of[1] != self._hist_version:
                    if self._ori_hist is not None:
                        t.set_ori_histogram(self._ori_hist)
                    if self._pos_hist is not None:
                        t.set_pos_histogram(self._pos_hist)
                t._twin_of = (self._sd_version, self._hist_version)
                if t.image_dtype != self.image_dtype:
                    t.set_image_dtype(self.image_dtype)
        return [self] + self._twins[:n - 1]

    def set_image_dtype(self, dtype: torch.dtype):
        """torch.float32 (reference contract, default) or torch.uint8 (pixels before ToTensor's /255; the stem divides)."""
        assert dtype in (torch.float32, torch.uint8)
        self._ck(self.lib.spef_set_image_dtype(self._h, 1 if dtype == torch.uint8 else 0))
        self.image_dtype = dtype

    def set_host_pack(self, on: bool):
        """Packed upload of float host images in eval_submit_host (spef_set_host_pack): host threads round the pixels to BF16 -- what the
        stem does first anyway -- so half the bytes cross the bus; results are bit-identical.  Default on for the BF16 engine."""
        self._ck(self.lib.spef_set_host_pack(self._h, 1 if on else 0))
        for t in getattr(self, "_twins", []):
            t.set_host_pack(on)

    def host_pack_info(self):
        """(active, host threads) of the packed upload for the next eval_submit_host."""
        import ctypes as C
        a, t = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.spef_host_pack_info(self._h, C.byref(a), C.byref(t), None))
        return bool(a.value), int(t.value)

    def host_pack_stats(self):
        """{packed fraction of the last submit, conversion rate, copy rate} as the context measured them (bytes / s)."""
        import ctypes as C
        st = np.zeros(3, np.float64)
        self._ck(self.lib.spef_host_pack_info(self._h, None, None, st.ctypes.data))
        return {"packed_fraction": float(st[0]), "convert_GBps": float(st[1]) / 1e9, "copy_GBps": float(st[2]) / 1e9}

    # ---- helpers ---------------------------------------------------------------------------------
    def _dev_f32(self, x, shape=None) -> torch.Tensor:
        t = torch.as_tensor(x)
        t = t.to(self.device, torch.float32, non_blocking=True).contiguous()
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), f"expected shape {shape}, got {tuple(t.shape)}"
        return t

    def _dev_img(self, images: torch.Tensor) -> torch.Tensor:
        if self.image_dtype == torch.uint8:
            assert images.dtype == torch.uint8, "engine is set to uint8 images"
            return images.to(self.device, non_blocking=True).contiguous()
        return self._dev_f32(images)

    def _empty(self, *shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def _check_images(self, images: torch.Tensor) -> int:
        if images.dim() != 4 or images.shape[1] != 3 or images.shape[2] != self.img_h or images.shape[3] != self.img_w:
            raise ValueError(f"images must be [B,3,{self.img_h},{self.img_w}], got {tuple(images.shape)}")
        if images.shape[0] > self.max_batch:
            raise ValueError(f"batch {images.shape[0]} exceeds the engine's max_batch {self.max_batch}")
        return images.shape[0]

    # ---- camera frames -> image tensor (input side of the path) ----------------------------------
    def resize_frames(self, frames: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        """SPEDataset's transform on a batch of decoded frames: Image.convert("RGB") -> Resize((img_h, img_w)) -> ToTensor
        (src/data/utils.py:215-226, src/data/datasets/speed.py:59-62), bit-exact against torchvision + Pillow.
        frames: uint8 [B,H,W] (grey) or [B,H,W,3] (RGB, HWC); returns [B,3,img_h,img_w] on this device, uint8 (the pixels
        ToTensor divides by 255; feed to an engine set to uint8 images) or float32 (the reference tensor).  Default:
        the engine's image dtype."""
        if frames.dtype != torch.uint8 or frames.dim() not in (3, 4) or (frames.dim() == 4 and frames.shape[3] != 3):
            raise ValueError(f"frames must be uint8 [B,H,W] or [B,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
        out_dtype = out_dtype or self.image_dtype
        assert out_dtype in (torch.float32, torch.uint8)
        f = frames.to(self.device, non_blocking=True).contiguous()
        B, H, W = f.shape[0], f.shape[1], f.shape[2]
        out = self._empty(B, 3, self.img_h, self.img_w, dtype=out_dtype)
        self._ck(self.lib.spef_resize_frames(self._h, ptr(f), B, H, W, 1 if f.dim() == 3 else 3, ptr(out),
                                             1 if out_dtype == torch.uint8 else 0, _stream(self.device)))
        return out

    # ---- network ---------------------------------------------------------------------------------
    def forward(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """ModelWrapper.forward: images [B,3,H,W] f32 on this device -> (ori logits [B,n_ori], pos [B,n_pos])."""
        B = self._check_images(images)
        x = self._dev_img(images)
        ori, pos = self._empty(B, self.n_ori), self._empty(B, self.n_pos)
        self._ck(self.lib.spef_forward(self._h, ptr(x), B, ptr(ori), ptr(pos), _stream(self.device)))
        return ori, pos

    def forward_timed(self, images: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, np.ndarray]:
        B = self._check_images(images)
        x = self._dev_img(images)
        ori, pos = self._empty(B, self.n_ori), self._empty(B, self.n_pos)
        ms = np.zeros(self.num_layers(), np.float32)
        self._ck(self.lib.spef_forward_timed(self._h, ptr(x), B, ptr(ori), ptr(pos), ms.ctypes.data, _stream(self.device)))
        return ori, pos, ms

    def num_layers(self) -> int:
        return self.lib.spef_num_layers(self._h)

    def layer_info(self, i: int) -> dict:
        v = [C.c_int32() for _ in range(10)]
        self._ck(self.lib.spef_layer_info(self._h, i, *[C.byref(x) for x in v]))
        names = ("kind", "cin", "cout", "hin", "win", "hout", "wout", "stride", "relu", "residual")
        return {n: x.value for n, x in zip(names, v)}

    def layer_forward(self, i: int, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Teacher-forced single layer.  x: stem -> [B,3,H,W] f32; otherwise NHWC [B,H,W,C] in the activation dtype
        (pool/head: [B,HW,C] / [B,C]).  Returns the layer output (NHWC, activation dtype; head: f32 [B,n_pad])."""
        info = self.layer_info(i)
        B = x.shape[0]
        x = x.to(self.device).contiguous()
        if info["kind"] == _ffi.KIND_HEAD:
            out = self._empty(B, info["cout"])
        elif info["kind"] == _ffi.KIND_POOL:
            out = self._empty(B, info["cout"], dtype=self.act_dtype)
        else:
            out = self._empty(B, info["hout"], info["wout"], info["cout"], dtype=self.act_dtype)
        want = torch.float32 if info["kind"] == _ffi.KIND_STEM else self.act_dtype
        assert x.dtype == want, f"layer {i} expects {want}, got {x.dtype}"
        r = None
        if residual is not None:
            r = residual.to(self.device).contiguous()
            assert r.dtype == self.act_dtype and r.shape == out.shape
        self._ck(self.lib.spef_layer_forward(self._h, i, ptr(x), ptr(r), ptr(out), B, _stream(self.device)))
        return out

    # ---- fused InvertedResidual blocks ----------------------------------------------------------------
    def num_blocks(self) -> int:
        return self.lib.spef_num_blocks(self._h)

    def block_info(self, i: int) -> dict:
        v = [C.c_int32() for _ in range(8)]
        self._ck(self.lib.spef_block_info(self._h, i, *[C.byref(x) for x in v]))
        names = ("first_layer", "n_layers", "fused", "tile_h", "tile_w", "groups", "w_stages", "resident")
        return {n: x.value for n, x in zip(names, v)}

    def fp32_hidden_blocks(self):
        """Feature indices (1..17, the reference's `features.features.{i}`) of the blocks that run as a channel-lane fused kernel
        in the current configuration: their hidden tensor stays FP32 on the SM (a rounding point the oracle needs to know)."""
        blocks = {i + 1 for i in range(self.num_blocks()) if self.block_info(i)["fused"] == 2}
        if self.stem_fusion_active():
            blocks.add(0)   # 0 = the stem: its output stays FP32 inside the stem + first-block kernel
        return blocks

    def stem_fusion_active(self) -> bool:
        """True when the stem conv runs inside the first block's kernel (spef_stem_fusion_active)."""
        return bool(self.lib.spef_stem_fusion_active(self._h))

    def pool_fusion_active(self) -> bool:
        """True when the last 1x1 conv and the global average pool run as one kernel (spef_pool_fusion_active)."""
        return bool(self.lib.spef_pool_fusion_active(self._h))

    def set_stem_fusion(self, on: bool):
        """True (default): stem + first InvertedResidual block as one kernel; False: separate stem launch."""
        self._ck(self.lib.spef_set_stem_fusion(self._h, 1 if on else 0))

    def stem_block_forward(self, images: torch.Tensor) -> torch.Tensor:
        """Teacher-forced stem + first block: images [B,3,H,W] (float32, or uint8 after set_image_dtype) -> NHWC bf16 block output."""
        bi = self.block_info(0)
        last = self.layer_info(bi["first_layer"] + bi["n_layers"] - 1)
        images = images.to(self.device).contiguous()
        out = self._empty(images.shape[0], last["hout"], last["wout"], last["cout"], dtype=torch.bfloat16)
        self._ck(self.lib.spef_stem_block_forward(self._h, ptr(images), ptr(out), images.shape[0], _stream(self.device)))
        return out

    def set_fusion(self, on: bool):
        """True (default): InvertedResidual blocks run as one fused kernel each; False: per-layer kernels."""
        self._ck(self.lib.spef_set_fusion(self._h, 1 if on else 0))

    def block_forward(self, i: int, x: torch.Tensor) -> torch.Tensor:
        """Teacher-forced fused block: x NHWC [B,H,W,Cin] bf16 -> [B,Ho,Wo,Cout] bf16."""
        bi = self.block_info(i)
        last = self.layer_info(bi["first_layer"] + bi["n_layers"] - 1)
        x = x.to(self.device).contiguous()
        assert x.dtype == torch.bfloat16
        out = self._empty(x.shape[0], last["hout"], last["wout"], last["cout"], dtype=torch.bfloat16)
        self._ck(self.lib.spef_block_forward(self._h, i, ptr(x), ptr(out), x.shape[0], _stream(self.device)))
        return out

    # ---- post-processing (device tensors) --------------------------------------------------------
    def decode_ori(self, x: torch.Tensor, is_logits: bool, want_soft=False, want_hinv=False, want_argmax=False):
        x = self._dev_f32(x)
        B, n = x.shape
        soft = self._empty(B, n) if (want_soft and is_logits) else None
        quat = self._empty(B, 4)
        hinv = self._empty(B, 4, 4) if want_hinv else None
        amax = self._empty(B, dtype=torch.int32) if want_argmax else None
        flags = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._ck(self.lib.spef_decode_ori(self._h, ptr(x), B, n, int(is_logits), ptr(soft), ptr(quat), ptr(hinv),
                                          ptr(amax), ptr(flags), _stream(self.device)))
        return {"soft": soft if is_logits else x, "quat": quat, "hinv": hinv, "argmax": amax, "flags": flags}

    def decode_pos(self, x: torch.Tensor, is_logits: bool, want_soft=False):
        x = self._dev_f32(x)
        B, n = x.shape
        soft = self._empty(B, n) if (want_soft and is_logits) else None
        pos = self._empty(B, 3)
        flags = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._ck(self.lib.spef_decode_pos(self._h, ptr(x), B, n, int(is_logits), ptr(soft), ptr(pos), ptr(flags),
                                          _stream(self.device)))
        return {"soft": soft if is_logits else x, "pos": pos, "flags": flags}

    def encode_ori(self, quat: torch.Tensor, variance: float, masked: Optional[torch.Tensor] = None):
        """Batch of true orientations [B,4] -> soft-classification pdfs [B,n_bins] (spef_encode_ori); masked: uint8 [n] or None."""
        q = torch.as_tensor(quat).to(self.device, torch.float64).contiguous()   # labels are float64 in the reference's datasets
        B = q.shape[0]
        out = self._empty(B, self.ori_hist_n)
        flags = torch.zeros(B, dtype=torch.int32, device=self.device)
        m = None if masked is None else masked.to(self.device, torch.uint8).contiguous()
        self._ck(self.lib.spef_encode_ori(self._h, ptr(q), B, self.ori_hist_n, float(variance), ptr(m), ptr(out), ptr(flags), _stream(self.device)))
        return out, flags

    def encode_pos(self, pos: torch.Tensor, variance: float):
        t = torch.as_tensor(pos).to(self.device, torch.float64).contiguous()
        B = t.shape[0]
        out = self._empty(B, self.pos_hist_n)
        flags = torch.zeros(B, dtype=torch.int32, device=self.device)
        self._ck(self.lib.spef_encode_pos(self._h, ptr(t), B, self.pos_hist_n, float(variance), ptr(out), ptr(flags), _stream(self.device)))
        return out, flags

    def error_stats(self, x: torch.Tensor, column: int = 0) -> Dict[str, float]:
        """mean / population std / median / median absolute deviation of x[:, column] (or of a 1-D x) on the device."""
        x = self._dev_f32(x)
        stride = 1 if x.dim() == 1 else x.shape[1]
        view = x if x.dim() == 1 else x[:, column]
        out = np.zeros(4, np.float64)
        self._ck(self.lib.spef_error_stats(self._h, view.data_ptr(), stride, x.shape[0], out.ctypes.data, _stream(self.device)))
        return {"mean": float(out[0]), "std": float(out[1]), "median": float(out[2]), "mad": float(out[3])}

    def score(self, quat_pred, pos_pred, quat_true, pos_true, sums: Optional[torch.Tensor] = None, want_per_image=False):
        qp, tp = self._dev_f32(quat_pred), self._dev_f32(pos_pred)
        qt, tt = self._dev_f32(quat_true), self._dev_f32(pos_true)
        B = qp.shape[0]
        if sums is None:
            sums = torch.zeros(8, dtype=torch.float64, device=self.device)
        per = self._empty(B, 2) if want_per_image else None
        self._ck(self.lib.spef_score(self._h, ptr(qp), ptr(tp), ptr(qt), ptr(tt), B, ptr(sums), ptr(per), _stream(self.device)))
        return sums, per

    def _host_img(self, images: torch.Tensor) -> torch.Tensor:
        """Host image batch in the dtype the context was told to expect (spef_set_image_dtype): the *_host entry points copy
        B*3*H*W elements of THAT type, so a float batch handed to a uint8 context (or the reverse) must be converted here --
        uint8 pixels are ToTensor's input (value / 255), float images are ToTensor's output."""
        x = images.detach()
        if self.image_dtype == torch.uint8:
            if x.dtype != torch.uint8:
                if not torch.is_floating_point(x):
                    raise TypeError(f"uint8 engine: images must be uint8 pixels or float images in [0, 1], got {x.dtype}")
                x = (x.to("cpu", torch.float32) * 255.0).round().clamp_(0, 255).to(torch.uint8)
            return x.to("cpu").contiguous()
        if x.dtype == torch.uint8:
            x = x.to("cpu", torch.float32) / 255.0   # torchvision ToTensor
        return x.to("cpu", torch.float32).contiguous()

    # ---- fused predict ---------------------------------------------------------------------------
    def predict(self, images: torch.Tensor, want_soft=True, want_argmax=False) -> Dict[str, torch.Tensor]:
        """forward + softmax + decode on device tensors (spef_predict)."""
        B = self._check_images(images)
        x = self._dev_img(images)
        out = {"ori": self._empty(B, 4), "pos": self._empty(B, 3), "flags": torch.zeros(B, dtype=torch.int32, device=self.device)}
        if want_soft:
            out["ori_soft"] = self._empty(B, self.n_ori)
            if self.pos_classification:
                out["pos_soft"] = self._empty(B, self.n_pos)
        if want_argmax:
            out["argmax"] = self._empty(B, dtype=torch.int32)
        self._ck(self.lib.spef_predict(self._h, ptr(x), B, ptr(out.get("ori_soft")), ptr(out["ori"]), ptr(out.get("pos_soft")),
                                       ptr(out["pos"]), ptr(out.get("argmax")), ptr(out["flags"]), _stream(self.device)))
        return out

    def predict_host(self, images: torch.Tensor, want_soft=True, want_argmax=False) -> Dict[str, np.ndarray]:
        """Host buffers in, host buffers out (spef_predict_host): the call SPETorch.predict maps to."""
        B = self._check_images(images)
        x = self._host_img(images)
        out = {"ori": np.empty((B, 4), np.float32), "pos": np.empty((B, 3), np.float32), "flags": np.zeros(B, np.uint32)}
        if want_soft:
            out["ori_soft"] = np.empty((B, self.n_ori), np.float32)
            if self.pos_classification:
                out["pos_soft"] = np.empty((B, self.n_pos), np.float32)
        if want_argmax:
            out["argmax"] = np.empty(B, np.int32)
        self._ck(self.lib.spef_predict_host(self._h, x.data_ptr(), B, ptr(out.get("ori_soft")), ptr(out["ori"]),
                                            ptr(out.get("pos_soft")), ptr(out["pos"]), ptr(out.get("argmax")),
                                            ptr(out["flags"]), _stream(self.device)))
        return out

    # ---- evaluation accumulators -----------------------------------------------------------------
    def eval_reset(self):
        self._ck(self.lib.spef_eval_reset(self._h, _stream(self.device)))

    def eval_batch(self, images: torch.Tensor, quat_true, pos_true, want_per_image=False):
        B = self._check_images(images)
        if images.device.type == "cpu":
            x = self._host_img(images)
            qt = np.ascontiguousarray(torch.as_tensor(quat_true).cpu().numpy(), np.float32)
            tt = np.ascontiguousarray(torch.as_tensor(pos_true).cpu().numpy(), np.float32)
            per = np.empty((B, 2), np.float32) if want_per_image else None
            self._ck(self.lib.spef_eval_batch_host(self._h, x.data_ptr(), qt.ctypes.data, tt.ctypes.data, B, ptr(per),
                                                   _stream(self.device)))
            return per
        x, qt, tt = self._dev_img(images), self._dev_f32(quat_true, (B, 4)), self._dev_f32(pos_true, (B, 3))
        per = self._empty(B, 2) if want_per_image else None
        self._ck(self.lib.spef_eval_batch(self._h, ptr(x), ptr(qt), ptr(tt), B, ptr(per), _stream(self.device)))
        return per

    def eval_submit_host(self, images: torch.Tensor, quat_true: torch.Tensor, pos_true: torch.Tensor,
                         per_image_out: Optional[torch.Tensor] = None):
        """Pipelined batch step (spef_eval_submit_host): H2D of this batch overlaps the compute of the previous one.
        All tensors are CPU float32, contiguous (pinned for true overlap) and must stay alive until eval_wait()."""
        B = self._check_images(images)
        assert images.device.type == "cpu" and images.dtype == self.image_dtype and images.is_contiguous()
        for t in (quat_true, pos_true):
            assert t.device.type == "cpu" and t.dtype == torch.float32 and t.is_contiguous()
        assert quat_true.shape == (B, 4) and pos_true.shape == (B, 3)
        if per_image_out is not None:
            assert per_image_out.device.type == "cpu" and per_image_out.dtype == torch.float32 and per_image_out.shape == (B, 2)
        self._ck(self.lib.spef_eval_submit_host(self._h, images.data_ptr(), quat_true.data_ptr(), pos_true.data_ptr(), B,
                                                ptr(per_image_out), _stream(self.device)))

    def eval_wait(self):
        self._ck(self.lib.spef_eval_wait(self._h, _stream(self.device)))

    def eval_read(self) -> np.ndarray:
        s = np.zeros(8, np.float64)
        self._ck(self.lib.spef_eval_read(self._h, s.ctypes.data, _stream(self.device)))
        return s

    def eval_sums_tensor(self) -> torch.Tensor:
        """The 8 device accumulators as a torch view-free copy target for collectives: returns a fresh
        tensor holding the current sums (float64 [8])."""
        return torch.from_numpy(self.eval_read()).to(self.device)

    # ---- host-buffer post-processing (what SPEUtils binds to) ------------------------------------
    def decode_ori_host(self, x: np.ndarray, is_logits: bool, want_soft=False, want_hinv=False, want_argmax=False):
        x = np.ascontiguousarray(x, np.float32)
        B, n = x.shape
        soft = np.empty((B, n), np.float32) if (want_soft and is_logits) else None
        quat = np.empty((B, 4), np.float32)
        hinv = np.empty((B, 4, 4), np.float32) if want_hinv else None
        amax = np.empty(B, np.int32) if want_argmax else None
        flags = np.zeros(B, np.uint32)
        self._ck(self.lib.spef_decode_ori_host(self._h, x.ctypes.data, B, n, int(is_logits), ptr(soft), ptr(quat), ptr(hinv),
                                               ptr(amax), ptr(flags), _stream(self.device)))
        return {"soft": soft if is_logits else x, "quat": quat, "hinv": hinv, "argmax": amax, "flags": flags}

    def decode_pos_host(self, x: np.ndarray, is_logits: bool, want_soft=False):
        x = np.ascontiguousarray(x, np.float32)
        B, n = x.shape
        soft = np.empty((B, n), np.float32) if (want_soft and is_logits) else None
        pos = np.empty((B, 3), np.float32)
        flags = np.zeros(B, np.uint32)
        self._ck(self.lib.spef_decode_pos_host(self._h, x.ctypes.data, B, n, int(is_logits), ptr(soft), ptr(pos), ptr(flags),
                                               _stream(self.device)))
        return {"soft": soft if is_logits else x, "pos": pos, "flags": flags}

    def score_host(self, quat_pred, pos_pred, quat_true, pos_true, want_per_image=False):
        qp, tp = np.ascontiguousarray(quat_pred, np.float32), np.ascontiguousarray(pos_pred, np.float32)
        qt, tt = np.ascontiguousarray(quat_true, np.float32), np.ascontiguousarray(pos_true, np.float32)
        B = qp.shape[0]
        assert qp.shape == (B, 4) and qt.shape == (B, 4) and tp.shape == (B, 3) and tt.shape == (B, 3)
        sums = np.zeros(8, np.float64)
        per = np.empty((B, 2), np.float32) if want_per_image else None
        self._ck(self.lib.spef_score_host(self._h, qp.ctypes.data, tp.ctypes.data, qt.ctypes.data, tt.ctypes.data, B,
                                          sums.ctypes.data, ptr(per), _stream(self.device)))
        return sums, per

    # ---- temporal --------------------------------------------------------------------------------
    def temporal_reset(self, n_streams: int = 1):
        self._ck(self.lib.spef_temporal_reset(self._h, n_streams, _stream(self.device)))
        self._t_streams = n_streams

    def _temporal_out(self, S: int):
        t = {
            "still_ori_soft": self._empty(S, self.n_ori), "still_pos_soft": self._empty(S, self.n_pos),
            "still_quat": self._empty(S, 4), "still_pos": self._empty(S, 3),
            "video_ori_soft": self._empty(S, self.n_ori), "video_pos_soft": self._empty(S, self.n_pos),
            "video_quat": self._empty(S, 4), "video_pos": self._empty(S, 3),
            "ori_distance": self._empty(S), "pos_distance": self._empty(S),
            "flags": torch.zeros(S, dtype=torch.int32, device=self.device),
        }
        st = SpefTemporalOut(*[t[n].data_ptr() for n, _ in SpefTemporalOut._fields_])
        return t, st

    def temporal_step_logits(self, ori_logits, pos_logits) -> Dict[str, torch.Tensor]:
        o, p = self._dev_f32(ori_logits), self._dev_f32(pos_logits)
        S = o.shape[0]
        t, st = self._temporal_out(S)
        self._ck(self.lib.spef_temporal_step_logits(self._h, ptr(o), ptr(p), S, C.byref(st), _stream(self.device)))
        return t

    def _temporal_slots(self, S: int):
        """Persistent device buffers of the frame step for S streams: the frame staging tensor and ONE flat output buffer the
        spef_temporal_out pointers are carved from.  Stable pointers are what lets libspef_b200 replay the step as a CUDA graph
        (spef_temporal_step captures a call signature it sees twice); the caller gets a clone of the flat buffer, so results
        of earlier frames stay valid."""
        key = (S, self.image_dtype)
        slot = self._tslots.get(key) if hasattr(self, "_tslots") else None
        if slot is None:
            if not hasattr(self, "_tslots"):
                self._tslots = {}
            sizes = [(n, S * w) for n, w in (("still_ori_soft", self.n_ori), ("still_pos_soft", self.n_pos), ("still_quat", 4), ("still_pos", 3),
                                             ("video_ori_soft", self.n_ori), ("video_pos_soft", self.n_pos), ("video_quat", 4), ("video_pos", 3),
                                             ("ori_distance", 1), ("pos_distance", 1), ("flags", 1))]
            offs, total = {}, 0
            for n, sz in sizes:
                offs[n] = (total, sz)
                total += (sz + 3) // 4 * 4           # 16-byte aligned slices
            flat = torch.zeros(total, dtype=torch.float32, device=self.device)
            frame = torch.empty((S, 3, self.img_h, self.img_w), dtype=self.image_dtype, device=self.device)
            st = SpefTemporalOut(*[flat.data_ptr() + 4 * offs[n][0] for n, _ in SpefTemporalOut._fields_])
            slot = self._tslots[key] = (frame, flat, offs, st)
        return slot

    def temporal_step(self, images: torch.Tensor, apply_filter: bool = True) -> Dict[str, torch.Tensor]:
        S = self._check_images(images)
        frame, flat, offs, st = self._temporal_slots(S)
        x = images.detach()
        if x.dtype != self.image_dtype:
            x = self._dev_img(x)
        frame.copy_(x, non_blocking=True)
        self._ck(self.lib.spef_temporal_step(self._h, ptr(frame), S, int(apply_filter), C.byref(st), _stream(self.device)))
        res = flat.clone()
        shapes = {"still_ori_soft": (S, self.n_ori), "still_pos_soft": (S, self.n_pos), "still_quat": (S, 4), "still_pos": (S, 3),
                  "video_ori_soft": (S, self.n_ori), "video_pos_soft": (S, self.n_pos), "video_quat": (S, 4), "video_pos": (S, 3),
                  "ori_distance": (S,), "pos_distance": (S,), "flags": (S,)}
        t = {}
        for n, (o, sz) in offs.items():
            v = res[o:o + sz].view(shapes[n])
            t[n] = v.view(torch.int32) if n == "flags" else v
        if not apply_filter:
            t = {k: v for k, v in t.items() if not k.startswith("video_") and not k.endswith("_distance")}
        return t

    def pdf_filter(self, cur: torch.Tensor, state: torch.Tensor, has_state: torch.Tensor, n_coef: float, alpha: float):
        """TemporalPDF.update_pdf on caller-owned device state; returns (filtered pdf [S,n], distance [S])."""
        cur = self._dev_f32(cur)
        S, n = cur.shape
        assert state.shape == (S, n) and state.dtype == torch.float32 and state.is_cuda
        assert has_state.shape == (S,) and has_state.dtype == torch.int32 and has_state.is_cuda
        out, dist = self._empty(S, n), self._empty(S)
        self._ck(self.lib.spef_pdf_filter(self._h, ptr(cur), S, n, ptr(state), ptr(has_state), float(n_coef), float(alpha),
                                          ptr(out), ptr(dist), _stream(self.device)))
        return out, dist

    # ---- introspection ---------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.spef_launch_count(self._h))

    def forward_cost(self, batch: int) -> Tuple[float, float]:
        b, f = C.c_double(), C.c_double()
        self._ck(self.lib.spef_forward_cost(self._h, batch, C.byref(b), C.byref(f)))
        return b.value, f.value
