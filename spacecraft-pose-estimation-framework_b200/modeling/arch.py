"""Mobile-URSONet topology and state_dict layout (reference: src/modeling/backbone/mobilenet_v2.py:232-271,
src/modeling/common/pytorch_layers.py:65-98, src/modeling/head/ursonet.py:10-33)."""
from __future__ import annotations

from typing import List, Tuple

# t (expand ratio), c (out channels), n (repeats), s (stride of the first repeat) -- mobilenet_v2.py:240-249
INVERTED_RESIDUAL_SETTINGS = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                              (6, 320, 1, 1)]
STEM_CHANNELS = 32
LAST_CHANNELS = 1280


def block_table() -> List[dict]:
    """One entry per InvertedResidual (feature index 1..17)."""
    out, cin, idx = [], STEM_CHANNELS, 1
    for t, c, n, s in INVERTED_RESIDUAL_SETTINGS:
        for i in range(n):
            stride = s if i == 0 else 1
            out.append(dict(idx=idx, cin=cin, cout=c, hidden=int(round(cin * t)), stride=stride, expand=(t != 1),
                            residual=(stride == 1 and cin == c)))
            cin = c
            idx += 1
    return out


def conv_layers() -> List[dict]:
    """The 52 ConvBnAct layers in execution order: prefix, kind ('stem'|'pw'|'dw'), conv weight shape, relu,
    residual (project conv of a residual block)."""
    L = [dict(prefix="features.features.0", kind="stem", shape=(STEM_CHANNELS, 3, 3, 3), relu=True, residual=False)]
    for b in block_table():
        p, j = f"features.features.{b['idx']}.conv", 0
        if b["expand"]:
            L.append(dict(prefix=f"{p}.{j}", kind="pw", shape=(b["hidden"], b["cin"], 1, 1), relu=True, residual=False))
            j += 1
        L.append(dict(prefix=f"{p}.{j}", kind="dw", shape=(b["hidden"], 1, 3, 3), relu=True, residual=False, stride=b["stride"]))
        j += 1
        L.append(dict(prefix=f"{p}.{j}", kind="pw", shape=(b["cout"], b["hidden"], 1, 1), relu=False, residual=b["residual"]))
    L.append(dict(prefix="features.features.18", kind="pw", shape=(LAST_CHANNELS, 320, 1, 1), relu=True, residual=False))
    return L


def state_dict_spec(n_ori: int, n_pos: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, role) for the 316 entries, in the reference's registration order.
    role in {'conv', 'bn_weight', 'bn_bias', 'bn_mean', 'bn_var', 'bn_count', 'linear_weight', 'linear_bias'}."""
    spec = []
    for l in conv_layers():
        c = l["shape"][0]
        spec.append((l["prefix"] + ".0.weight", tuple(l["shape"]), "conv"))
        spec.append((l["prefix"] + ".1.weight", (c,), "bn_weight"))
        spec.append((l["prefix"] + ".1.bias", (c,), "bn_bias"))
        spec.append((l["prefix"] + ".1.running_mean", (c,), "bn_mean"))
        spec.append((l["prefix"] + ".1.running_var", (c,), "bn_var"))
        spec.append((l["prefix"] + ".1.num_batches_tracked", (), "bn_count"))
    # URSONetHead registers pos before ori (head/ursonet.py:17-25)
    spec.append(("head.pos.0.weight", (n_pos, LAST_CHANNELS), "linear_weight"))
    spec.append(("head.pos.0.bias", (n_pos,), "linear_bias"))
    spec.append(("head.ori.1.weight", (n_ori, LAST_CHANNELS), "linear_weight"))
    spec.append(("head.ori.1.bias", (n_ori,), "linear_bias"))
    return spec


def output_hw(h: int, w: int) -> Tuple[int, int]:
    """Spatial size of the last feature map (five stride-2 stages, k3 p1: out = (in + 2 - 3)//2 + 1)."""
    for _ in range(5):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    return h, w
