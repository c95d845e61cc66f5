"""Developer diagnostic: per fused block, error vs the oracle and vs the per-layer kernel chain (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import spef_oracle as O
from spef_b200.tools import synthetic
from spef_b200.engine import Engine
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_network import _oracle_layer_io, nhwc

sd = synthetic.synthetic_state_dict(1728, 3)
B = int(os.environ.get("DBG_B", "3"))
x = synthetic.synthetic_images(B)
eng = Engine(240, 384, 1728, 3, False, "bf16", 8, None, 0)
eng.load_state_dict(sd)
ios = _oracle_layer_io(sd, x, True)
for bi in range(eng.num_blocks()):
    info = eng.block_info(bi)
    if not info["fused"]:
        print(bi, "not fused", info); continue
    first, last = info["first_layer"], info["first_layer"] + info["n_layers"] - 1
    inp = nhwc(ios[first][0], torch.bfloat16)
    want = ios[last][2].permute(0, 2, 3, 1).contiguous()
    try:
        got = eng.block_forward(bi, inp).float().cpu()
    except Exception as e:
        print(bi, "FAILED", e); break
    cur = inp
    for li in range(first, last + 1):
        cur = eng.layer_forward(li, cur, inp if eng.layer_info(li)["residual"] else None)
    chain = cur.float().cpu()
    scale = float(want.abs().max())
    msg = f"block {bi:2d} v{info['fused']} L{first}-{last} tile {info['tile_h']}x{info['tile_w']} ng{info['groups']} w{info['w_stages']} res{info['resident']} shape {tuple(got.shape)}:"
    for name, ref in (("oracle", want), ("chain", chain)):
        err = (got - ref).abs()
        ulp = torch.maximum(ref.abs(), torch.tensor(scale * 2 ** -8)) * 2 ** -7
        r = err / ulp
        worst, frac = float(r.max()), float((err > 0).float().mean())
        idx = np.unravel_index(int(r.argmax()), r.shape)
        nbad = int((r > 2).sum())
        msg += f"  vs {name}: worst {worst:.2f} ulp at {tuple(int(v) for v in idx)}, differ {frac:.4f}, >2ulp {nbad}, nan {int(torch.isnan(got).sum())}"
    print(msg, flush=True)
    if os.environ.get("DBG_MAP") and bi == int(os.environ["DBG_MAP"]):
        bad = ((got - chain).abs() > 0).any(dim=3)[0].numpy().astype(int)
        for row in bad: print("".join(".#"[v] for v in row))
