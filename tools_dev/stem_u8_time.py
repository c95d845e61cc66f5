"""Developer tool (GPU box): time of the fused stem + block 1 launch (layer slot 0) with float and with uint8 images, B = 256."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from spef_b200.engine import Engine
from spef_b200.tools import synthetic

sd = synthetic.synthetic_state_dict(1728, 3)
eng = Engine(240, 384, 1728, 3, False, "bf16", 256, "cuda:0")
eng.load_state_dict(sd)
x = synthetic.synthetic_images(8).repeat(32, 1, 1, 1).contiguous()
xu = (x * 255).round().to(torch.uint8).cuda()
xf = x.cuda()
for name, img, dt in (("f32", xf, torch.float32), ("u8", xu, torch.uint8)):
    eng.set_image_dtype(dt)
    ms = np.median(np.stack([eng.forward_timed(img)[2] for _ in range(9)]), axis=0)
    print(name, "slot 0: %.1f us, whole forward %.1f us" % (ms[0] * 1000, ms.sum() * 1000), flush=True)
