"""Developer tool: N forwards at B = 256 (for ncu launch lists / captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spef_b200.tools import synthetic
from spef_b200.engine import Engine
B = int(os.environ.get("DBG_B", "256"))
n = int(os.environ.get("DBG_N", "3"))
eng = Engine(240, 384, 1728, 3, False, "bf16", B, None, 0)
eng.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
from spef_b200.spe.classification_utils import OrientationSoftClassification
eng.set_ori_histogram(OrientationSoftClassification(12, 3, False).histogram)
x = synthetic.synthetic_images(32).repeat(B // 32, 1, 1, 1).cuda()
tg = synthetic.synthetic_targets(B, 2024)
qt, tt = torch.from_numpy(tg["ori"]).cuda(), torch.from_numpy(tg["pos"]).cuda()
eng.eval_reset()
for _ in range(n):
    eng.eval_batch(x, qt, tt)
torch.cuda.synchronize()
print("ok", eng.launch_count())
