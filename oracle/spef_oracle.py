"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SPEF pose-inference hot path.

A CPU restatement (NumPy + torch-CPU functional ops) of the reference algorithm
for the path named in BASELINE.json: Mobile-URSONet forward -> softmax ->
soft-classification decode -> pose error / ESA score -> temporal pdf filter.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module; the product package never does
(it fails loudly when the CUDA extension is missing).

Where the arithmetic really lives: the reference network is ``torch.nn`` modules
(third-party: PyTorch, reference pins 1.10 / 2.1, this image 2.11) and the
post-processing is NumPy + LAPACK (unpinned in the reference, 2.3.5 here).  The
restatement calls the same third-party primitives (``F.conv2d``, ``np.linalg``)
in the order the cited reference lines do.

PARITY PIN: the reference ships no tests or golden vectors.
This oracle is pinned against outputs of the *unmodified reference itself*, run in
the build container through ``oracle/ref_loader.py`` and frozen under
``tests/golden/*.npz`` by ``tests/golden/make_goldens.py``;
``tests/test_oracle_golden.py`` checks every function below against them.

All ``file:line`` citations are relative to the reference root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Architecture table  (src/modeling/backbone/mobilenet_v2.py:240-249, 252-264)
# --------------------------------------------------------------------------------------
IR_SETTINGS = [  # t, c, n, s
    (1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1),
]
BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, src/modeling/common/pytorch_layers.py:55-56


def block_table() -> List[dict]:
    """One dict per InvertedResidual, feature index 1..17 (mobilenet_v2.py:257-262)."""
    out, cin, idx = [], 32, 1
    for t, c, n, s in IR_SETTINGS:
        for i in range(n):
            stride = s if i == 0 else 1
            out.append(dict(idx=idx, cin=cin, cout=c, hidden=int(round(cin * t)), stride=stride, expand=(t != 1),
                            residual=(stride == 1 and cin == c)))  # pytorch_layers.py:71,73
            cin = c
            idx += 1
    return out


def _bn(sd, prefix):
    return (sd[prefix + ".weight"], sd[prefix + ".bias"], sd[prefix + ".running_mean"], sd[prefix + ".running_var"])


def _conv_bn_act(x, sd, prefix, stride, groups, act, rnd=None):
    """ConvBnAct in eval mode: Conv2d(bias=False) -> BatchNorm2d(running stats) -> ReLU
    (pytorch_layers.py:35-62; padding (k-1)//2 at :50-51; ReLU not ReLU6 at :59-60)."""
    w = sd[prefix + ".0.weight"]
    k = w.shape[-1]
    y = F.conv2d(x, w, None, stride, (k - 1) // 2, 1, groups)
    g, b, m, v = _bn(sd, prefix + ".1")
    y = F.batch_norm(y, m, v, g, b, False, 0.0, BN_EPS)
    if act:
        y = F.relu(y)
    return y


def forward_fp32(sd: Dict[str, torch.Tensor], images: torch.Tensor, return_features: bool = False):
    """ModelWrapper.forward = head(features(x))  (pytorch_layers.py:29-32).

    images: [B,3,H,W] float32 NCHW in [0,1].  Returns (ori_logits [B,n_ori], pos [B,n_pos]).
    With return_features=True also returns the list of activations after feature index 0..18.
    """
    feats = []
    with torch.no_grad():
        x = _conv_bn_act(images, sd, "features.features.0", 2, 1, True)  # mobilenet_v2.py:252-254
        feats.append(x)
        for blk in block_table():  # pytorch_layers.py:65-98
            p = f"features.features.{blk['idx']}.conv"
            y, j = x, 0
            if blk["expand"]:  # :77-79
                y = _conv_bn_act(y, sd, f"{p}.{j}", 1, 1, True)
                j += 1
            y = _conv_bn_act(y, sd, f"{p}.{j}", blk["stride"], blk["hidden"], True)  # :82-83
            j += 1
            y = _conv_bn_act(y, sd, f"{p}.{j}", 1, 1, False)  # :85-86 linear bottleneck
            x = x + y if blk["residual"] else y  # :93-98
            feats.append(x)
        x = _conv_bn_act(x, sd, "features.features.18", 1, 1, True)  # mobilenet_v2.py:264
        feats.append(x)
        f = x.mean([2, 3])  # head/ursonet.py:30
        ori = F.linear(f, sd["head.ori.1.weight"], sd["head.ori.1.bias"])  # :31 (Dropout = identity in eval)
        pos = F.linear(f, sd["head.pos.0.weight"], sd["head.pos.0.bias"])  # :32
    if return_features:
        return ori, pos, feats
    return ori, pos


# --------------------------------------------------------------------------------------
# BN-folded / fake-BF16 variant: same algorithm with the rounding points of the CUDA BF16 path
# (DESIGN.md "rounding points").  Used for per-kernel teacher-forced parity and end-to-end BF16
# parity; it is NOT a restatement of reference behaviour beyond forward_fp32.
# --------------------------------------------------------------------------------------
def fold_bn(w: torch.Tensor, g, b, m, v) -> Tuple[torch.Tensor, torch.Tensor]:
    """s = g/sqrt(v+eps); w' = w*s; b' = b - m*s   (SURVEY Appendix A.1), computed in float64."""
    s = g.double() / torch.sqrt(v.double() + BN_EPS)
    wf = (w.double() * s.view(-1, 1, 1, 1)).float()
    bf = (b.double() - m.double() * s).float()
    return wf, bf


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def folded_layers(sd) -> List[dict]:
    """Flat list of the 52 conv layers + meta, BN folded (float32)."""
    L = []

    def add(prefix, kind, stride, act, residual=False, block=None, role=None):
        w = sd[prefix + ".0.weight"]
        wf, bf = fold_bn(w, *_bn(sd, prefix + ".1"))
        L.append(dict(prefix=prefix, kind=kind, stride=stride, act=act, residual=residual, w=wf, b=bf, block=block, role=role or kind))

    add("features.features.0", "stem", 2, True)
    for blk in block_table():
        p = f"features.features.{blk['idx']}.conv"
        j = 0
        if blk["expand"]:
            add(f"{p}.{j}", "pw", 1, True, block=blk["idx"], role="expand")
            j += 1
        add(f"{p}.{j}", "dw", blk["stride"], True, block=blk["idx"], role="dw")
        j += 1
        add(f"{p}.{j}", "pw", 1, False, residual=blk["residual"], block=blk["idx"], role="project")
    add("features.features.18", "pw", 1, True)
    return L


def apply_layer(layer: dict, x: torch.Tensor, res: Optional[torch.Tensor], bf16: bool, round_out: bool = True) -> torch.Tensor:
    """One folded conv layer on NCHW float32 tensors, optionally with BF16 rounding points:
    pointwise and stem weights (and the stem's image taps) rounded to bf16 (tensor-core operands), depthwise weights kept
    f32, fp32 accumulate, +bias, ReLU, (+residual), output rounded to bf16 (round_out=False: the expand output of a block whose
    fused kernel keeps the hidden tensor in FP32 on the SM -- it never reaches HBM, so there is nothing to round for)."""
    w, b = layer["w"], layer["b"]
    if layer["kind"] in ("pw", "stem") and bf16:
        w = bf16_round(w)  # tensor-core operands (the stem runs as an implicit GEMM: image taps and weights in BF16)
    if layer["kind"] == "stem" and bf16:
        x = bf16_round(x)
    groups = w.shape[0] if layer["kind"] == "dw" else 1
    k = w.shape[-1]
    y = F.conv2d(x, w, b, layer["stride"], (k - 1) // 2, 1, groups)
    if layer["act"]:
        y = F.relu(y)
    if res is not None:
        y = res + y
    return bf16_round(y) if (bf16 and round_out) else y


def forward_folded(sd, images: torch.Tensor, bf16: bool = False, return_layers: bool = False, fp32_hidden_blocks=()):
    """BN-folded forward; bf16=True reproduces the CUDA BF16 path's rounding points.  fp32_hidden_blocks: feature indices
    (1..17) of the InvertedResidual blocks that run as a channel-lane fused kernel, whose hidden tensor (expand output) stays
    FP32 between the expand GEMM and the depthwise taps (Engine.fp32_hidden_blocks()); 0 stands for the stem fused into the first
    block's kernel (its output, the hidden tensor of that kernel, is not rounded either)."""
    fp32_hidden_blocks = set(fp32_hidden_blocks)
    outs = []
    with torch.no_grad():
        x = images
        block_in = None
        for layer in folded_layers(sd):
            first_of_block = layer["kind"] == "stem" or layer["prefix"].endswith("conv.0") or layer["prefix"].endswith(".18")
            if first_of_block:
                block_in = x
            res = block_in if layer["residual"] else None
            keep_fp32 = (layer["role"] == "expand" and layer["block"] in fp32_hidden_blocks) or (layer["kind"] == "stem" and 0 in fp32_hidden_blocks)
            x = apply_layer(layer, x, res, bf16, round_out=not keep_fp32)
            outs.append(x)
        f = x.mean([2, 3])
        wo, wp = sd["head.ori.1.weight"], sd["head.pos.0.weight"]
        if bf16:
            f, wo, wp = bf16_round(f), bf16_round(wo), bf16_round(wp)
        ori = F.linear(f, wo, sd["head.ori.1.bias"])
        pos = F.linear(f, wp, sd["head.pos.0.bias"])
    if return_layers:
        return ori, pos, outs
    return ori, pos


# --------------------------------------------------------------------------------------
# Post-processing  (src/spe/spe_utils.py, src/spe/classification_utils.py, src/spe/utils.py)
# --------------------------------------------------------------------------------------
def softmax(z: np.ndarray) -> np.ndarray:
    """Row-wise max-subtracted softmax, dtype preserved (spe_utils.py:75-76 / :77-79)."""
    e = np.exp(z - np.max(z, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def euler2quat(yaw: float, pitch: float, roll: float) -> np.ndarray:
    """Degrees, ZYX, scalar-first (src/spe/utils.py:211-230, gymbal_check irrelevant to the value)."""
    cy, sy = np.cos(np.deg2rad(yaw) / 2), np.sin(np.deg2rad(yaw) / 2)
    cp, sp = np.cos(np.deg2rad(pitch) / 2), np.sin(np.deg2rad(pitch) / 2)
    cr, sr = np.cos(np.deg2rad(roll) / 2), np.sin(np.deg2rad(roll) / 2)
    q = np.array([cy * cp * cr + sy * sp * sr, cy * cp * sr - sy * sp * cr,
                  cy * sp * cr + sy * cp * sr, sy * cp * cr - cy * sp * sr])
    return q / np.linalg.norm(q)


def ori_histogram(n_per_dim: int, delete_unused: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """classification_utils.py:39-83.  Returns (quaternion bins [n,4] f64, redundant flags [n^3] bool)."""
    min_lim, max_lim = np.array([-180, -90, -180]), np.array([180, 90, 180])
    g = np.linspace(0.0, 1.0, n_per_dim)
    grid = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    eul = grid * (max_lim - min_lim) + min_lim
    q = np.zeros((n_per_dim ** 3, 4))
    for i in range(q.shape[0]):
        q[i] = euler2quat(eul[i, 0], eul[i, 1], eul[i, 2])
    boundary = np.logical_or(eul[:, 0] == max_lim[0], eul[:, 2] == max_lim[2])
    gimbal = np.logical_and(np.abs(eul[:, 1]) == max_lim[1], eul[:, 0] != min_lim[0])
    red = np.logical_or(boundary, gimbal)
    if delete_unused:
        q = q[~red]
    return q, red


def ori_encode(q_true: np.ndarray, hist: np.ndarray, red: np.ndarray, n_per_dim: int, smooth: float,
               delete_unused: bool = False) -> np.ndarray:
    """classification_utils.py:85-111."""
    var = (smooth / n_per_dim) ** 2 / 12
    k = np.exp(-((2 * np.arccos(np.minimum(1.0, np.abs(np.sum(q_true * hist, axis=1)))) / np.pi) ** 2) / (2 * var))
    if not delete_unused:
        k[red] = 0
    p = k / np.sum(k)
    if np.any(np.isnan(p)):
        raise ValueError("NaN found in encoded orientation")
    return p.astype(np.float32)


def ori_decode(p: np.ndarray, hist: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """classification_utils.py:113-147: A = sum_b p_b q_b q_b^T (f64), dominant eigenvector via
    np.linalg.eig + argsort, renormalise, h_inv = inv(A); both cast to float32."""
    b = hist.reshape(-1, 4, 1) * hist.reshape(-1, 1, 4)  # :168-176
    a = np.sum(b * np.reshape(p, (-1, 1, 1)), axis=0)
    if np.any(np.isnan(a)):
        raise ValueError("Error during orientation decoding")
    s, v = np.linalg.eig(a)
    q = v[:, np.argsort(s)[-1]]
    q = q / np.linalg.norm(q)
    return np.real(q).astype(np.float32), np.linalg.inv(a).astype(np.float32)


def ori_decode_batch(pb: np.ndarray, hist: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """classification_utils.py:149-166."""
    qs = np.zeros((pb.shape[0], 4), np.float32)
    hs = np.zeros((pb.shape[0], 4, 4), np.float32)
    for i in range(pb.shape[0]):
        qs[i], hs[i] = ori_decode(pb[i], hist)
    return qs, hs


POS_MIN = np.array([-16, -12, -2])  # spe_utils.py:51-53
POS_MAX = np.array([16, 12, 40])


def pos_histogram(n_per_dim: int, min_lim=POS_MIN, max_lim=POS_MAX) -> np.ndarray:
    """classification_utils.py:201-216."""
    g = np.linspace(0.0, 1.0, n_per_dim)
    grid = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    return grid * (max_lim - min_lim) + min_lim


def pos_encode(t: np.ndarray, hist: np.ndarray, n_per_dim: int, smooth: float) -> np.ndarray:
    """classification_utils.py:218-240."""
    var = (smooth / n_per_dim) ** 2 / 12
    k = np.exp(-np.sum((t - hist) ** 2, axis=1) / (2 * var))
    p = k / np.sum(k)
    if np.any(np.isnan(p)):
        raise ValueError("NaN found in encoded position")
    return p.astype(np.float32)


def pos_decode(p: np.ndarray, hist: np.ndarray) -> np.ndarray:
    """classification_utils.py:242-267."""
    if np.sum(p) == 0:
        raise ValueError("Encoded position vector sum is zero, cannot decode.")
    t = np.sum(hist * np.reshape(p, (-1, 1)), axis=0) / np.sum(p)
    if np.any(np.isnan(t)):
        raise ValueError("Error during position decoding, NaN found in decoded position.")
    return t.astype(np.float32)


def pos_decode_batch(pb: np.ndarray, hist: np.ndarray) -> np.ndarray:
    """classification_utils.py:269-285."""
    return np.stack([pos_decode(pb[i], hist) for i in range(pb.shape[0])]).astype(np.float32)


def get_score(true_pose: dict, pred_pose: dict) -> dict:
    """spe_utils.py:104-159 (dtype follows the inputs: float32 in evaluation()).  The `> 1.01`
    ValueError at :137-138 is dead code in the reference (`True in arr` compares floats with
    True), so the observed behaviour -- clamp only -- is what is restated."""
    qt, tt, qp, tp = true_pose["ori"], true_pose["pos"], pred_pose["ori"], pred_pose["pos"]
    e_t = np.linalg.norm(tt - tp, axis=1)
    e_tn = e_t / np.linalg.norm(tt, axis=1)
    c = np.abs(np.sum(qp * qt, axis=1, keepdims=True))
    if True in c[c > 1.01]:  # kept verbatim in meaning: fires only if an element == 1.0 exactly -> never
        raise ValueError("Intermediate sum issue due to error in model prediction (orientation)")
    c[c > 1] = 1
    ori = np.mean(2 * np.arccos(c))
    pos = np.mean(e_tn)
    return {"esa_score": ori + pos, "ori_score": ori, "pos_score": pos,
            "ori_error": ori * 180 / np.pi, "pos_error": np.mean(e_t)}


def per_image_errors(true_pose: dict, pred_pose: dict) -> Tuple[np.ndarray, np.ndarray]:
    """src/tools/evaluation.py:82-85: (ori error in degrees, pos error in metres) per image."""
    pos = np.linalg.norm(true_pose["pos"] - pred_pose["pos"], axis=1)
    c = np.abs(np.sum(pred_pose["ori"] * true_pose["ori"], axis=1, keepdims=True))
    c[c > 1] = 1
    return (2 * np.arccos(c) * 180 / np.pi).reshape(-1), pos


def mad(data) -> float:
    """src/tools/evaluation.py:16-32."""
    med = np.median(data)
    return np.median(np.abs(np.array(data) - med)).tolist()


class RunningMean:
    """Batch-weighted running mean = AverageMeter (src/tools/utils.py:67-104)."""

    def __init__(self, keys):
        self.sum = {k: 0.0 for k in keys}
        self.count = 0

    def update(self, values: dict, n: int):
        for k, v in values.items():
            self.sum[k] += v * n
        self.count += n

    def get(self, k):
        return self.sum[k] / self.count


def evaluate(batches, predict_fn) -> Tuple[dict, dict]:
    """One phase of evaluation() (src/tools/evaluation.py:59-99).  `batches` yields
    (images [B,3,H,W] torch, targets {'ori': [B,4], 'pos': [B,3]} numpy f32); `predict_fn`
    maps images -> pose dict with 'ori', 'pos'.  Returns (score, error) dicts of floats."""
    keys = ("esa_score", "ori_score", "pos_score", "ori_error", "pos_error")
    avg = RunningMean(keys)
    e_ori, e_pos = [], []
    for images, targets in batches:
        pose = predict_fn(images)
        avg.update(get_score(targets, pose), images.shape[0])
        eo, ep = per_image_errors(targets, pose)
        e_ori.extend(eo)
        e_pos.extend(ep)
    score = {"ori": avg.get("ori_score"), "pos": avg.get("pos_score"), "esa": avg.get("esa_score")}
    error = {"ori": avg.get("ori_error"), "pos": avg.get("pos_error"),
             "ori_std": np.std(e_ori).tolist(), "pos_std": np.std(e_pos).tolist(),
             "ori_mad": mad(e_ori), "pos_mad": mad(e_pos)}
    return score, error


# --------------------------------------------------------------------------------------
# Temporal filter  (src/temporal/pdf_compare.py:94-133, src/temporal/inference.py:114-195)
# --------------------------------------------------------------------------------------
class TemporalPDF:
    """pdf_compare.py: adaptive recursive blend with the 'l2' metric (:50-52, :80-92, :94-133)."""

    def __init__(self, n: float, alpha: float):
        self.n, self.alpha, self.previous_pdf = n, alpha, None

    def reset(self):
        self.previous_pdf = None

    def update_pdf(self, cur: np.ndarray):
        cur = cur / np.sum(cur)
        if self.previous_pdf is None:
            self.previous_pdf = cur
            return cur, 0.0
        p1, p2 = cur / np.sum(cur), self.previous_pdf / np.sum(self.previous_pdf)
        d = np.linalg.norm(p1 - p2)
        w = np.clip(np.exp(-self.alpha * d), 0.0, 1.0)
        upd = w * self.n * cur + (1 - w) * self.previous_pdf
        upd = upd / np.sum(upd)
        self.previous_pdf = upd
        return upd, d


def sign_continuity(prev: Optional[np.ndarray], q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """inference.py:136-144 / :173-180.  Returns (new prev, possibly flipped q)."""
    if prev is None:
        return q, q
    dot = np.dot(prev, q)
    if dot < 0:
        q = -q
    if np.abs(dot) > 0.5:
        prev = q
    return prev, q


class TemporalInference:
    """Inference.predict(image, 'Adaptative') after the network: still decode + filtered decode
    (inference.py:38-39 parameters, :131-180 flow).  Operates on raw logits of one frame."""

    def __init__(self, ori_hist: np.ndarray, pos_hist: np.ndarray):
        self.ori_hist, self.pos_hist = ori_hist, pos_hist
        self.f_ori, self.f_pos = TemporalPDF(0.8, 16.49), TemporalPDF(0.5, 48.64)
        self.prev_still = self.prev_video = None

    def reset(self):
        self.prev_still = self.prev_video = None
        self.f_ori.reset()
        self.f_pos.reset()

    def step(self, ori_logits: np.ndarray, pos_logits: np.ndarray):
        ori_soft = softmax(ori_logits[None].astype(np.float32))[0]
        pos_soft = softmax(pos_logits[None].astype(np.float32))[0]
        q, _ = ori_decode(ori_soft, self.ori_hist)
        still = {"ori_soft": ori_soft, "pos_soft": pos_soft, "ori": q, "pos": pos_decode(pos_soft, self.pos_hist)}
        self.prev_still, still["ori"] = sign_continuity(self.prev_still, still["ori"])
        video = {}
        video["ori_soft"], video["ori_distance"] = self.f_ori.update_pdf(ori_soft)
        video["pos_soft"], video["pos_distance"] = self.f_pos.update_pdf(pos_soft)
        video["ori"], _ = ori_decode(video["ori_soft"], self.ori_hist)
        video["pos"] = pos_decode(video["pos_soft"], self.pos_hist)
        self.prev_video, video["ori"] = sign_continuity(self.prev_video, video["ori"])
        return still, video


def quat_angle_deg(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Sign-invariant angle between unit quaternions in float64, well conditioned near 0
    (2*atan2(|a-+b|, |a+-b|))."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    dm, dp = np.linalg.norm(a - b, axis=-1), np.linalg.norm(a + b, axis=-1)
    return np.degrees(2 * np.arctan2(np.minimum(dm, dp), np.maximum(dm, dp))) * 2


# --------------------------------------------------------------------------------------
# Input side of the path: SPEDataset.__getitem__ (src/data/utils.py:212-226) opens the frame with
# PIL, .convert("RGB"), then applies transforms.Compose([Resize(img_size), ToTensor()]) (src/data/datasets/speed.py:59-62).
# The arithmetic lives in third-party Pillow (unpinned by the reference; 12.2.0 here): Image.resize(BILINEAR) is a
# separable, antialiased triangle filter evaluated in 22-bit fixed point on 8-bit pixels, horizontal pass first, each
# pass rounded and clipped to 8 bits; torchvision's ToTensor is float32(u8) / 255.  Restated from Pillow's published
# algorithm (libImaging "Resample": coefficient pre-computation in float64, 8-bit-per-channel passes) and pinned against
# the real torchvision + Pillow pipeline by tests/golden/resize.npz (tests/golden/make_goldens.py: golden_resize).
# --------------------------------------------------------------------------------------
RESIZE_PRECISION_BITS = 32 - 8 - 2


def resize_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Triangle-filter taps for one axis: (bounds [out,2] = first input index / tap count, kk [out,ksize] int32)."""
    scale = float(in_size) / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    inv = 1.0 / fscale
    for o in range(out_size):
        center = (o + 0.5) * scale
        lo = max(int(center - support + 0.5), 0)
        hi = min(int(center + support + 0.5), in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        total = 0.0
        for j in range(n):
            t = abs((j + lo - center + 0.5) * inv)
            w[j] = 1.0 - t if t < 1.0 else 0.0
            total += w[j]
        if total != 0.0:
            w = w / total
        bounds[o] = (lo, n)
        for j in range(n):
            v = w[j] * (1 << RESIZE_PRECISION_BITS)
            kk[o, j] = int(-0.5 + v) if w[j] < 0 else int(0.5 + v)
    return bounds, kk


def _resize_pass(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One 8-bit pass along axis -1 of img [..., in] -> [..., out]."""
    out = np.empty(img.shape[:-1] + (bounds.shape[0],), np.uint8)
    src = img.astype(np.int64)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = (src[..., lo:lo + n] * kk[o, :n].astype(np.int64)).sum(-1) + (1 << (RESIZE_PRECISION_BITS - 1))
        out[..., o] = np.clip(acc >> RESIZE_PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_frames(frames: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """frames uint8 [B,H,W] (grey, replicated like .convert('RGB')) or [B,H,W,3] -> uint8 [B,3,out_h,out_w] (the bytes
    ToTensor divides by 255)."""
    f = np.asarray(frames)
    assert f.dtype == np.uint8
    if f.ndim == 3:
        f = np.repeat(f[..., None], 3, axis=-1)
    f = np.ascontiguousarray(f.transpose(0, 3, 1, 2))          # [B,3,H,W]
    hb, hk = resize_coeffs(f.shape[3], out_w)
    vb, vk = resize_coeffs(f.shape[2], out_h)
    h = _resize_pass(f, hb, hk)                                 # [B,3,H,out_w]
    v = _resize_pass(np.ascontiguousarray(h.transpose(0, 1, 3, 2)), vb, vk)   # [B,3,out_w,out_h]
    return np.ascontiguousarray(v.transpose(0, 1, 3, 2))


def to_tensor(u8: np.ndarray) -> np.ndarray:
    """torchvision ToTensor on 8-bit data: float32(u8) / 255 (float32 division)."""
    return u8.astype(np.float32) / np.float32(255.0)
