import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from spef_b200.tools import synthetic
from spef_b200.engine import Engine
from oracle import spef_oracle as O
sd = synthetic.synthetic_state_dict(1728, 3)
eng = Engine(240, 384, 1728, 3, False, "bf16", 8, None, 0)
eng.load_state_dict(sd)
layers = O.folded_layers(sd)
if os.environ.get("DBG_U8"):
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 3, 240, 384), generator=g, dtype=torch.uint8)
    x = u8.float().div(255)
    eng.set_image_dtype(torch.uint8)
    got = eng.layer_forward(0, u8.cuda()).float().cpu() if False else None
    o, p = eng.forward(u8)
    print("u8 forward ok", float(o.abs().max()))
else:
    x = synthetic.synthetic_images(2)
    got = eng.layer_forward(0, x).float().cpu()
    want = O.apply_layer(layers[0], x, None, True).permute(0, 2, 3, 1)
    print("max err", float((got - want).abs().max()), "scale", float(want.abs().max()))
