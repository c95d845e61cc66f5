"""Developer tool: run the forward at B = 256 a few times with SPEF_FB_TRACE=<block> set; the library dumps the CTA-0 trace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spef_b200.tools import synthetic
from spef_b200.engine import Engine
B = int(os.environ.get("DBG_B", "256"))
eng = Engine(240, 384, 1728, 3, False, "bf16", B, None, 0)
eng.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
x = synthetic.synthetic_images(32).repeat(B // 32, 1, 1, 1).cuda()
for _ in range(6):
    eng.forward(x)
torch.cuda.synchronize()
