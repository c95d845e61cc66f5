"""Developer tool: one small invocation of the whole hot path (B = 4: forward with every fused block, softmax + decode, score, and one
temporal frame step on Mobile-URSONet+) for compute-sanitizer:
    compute-sanitizer --tool racecheck  python tools_dev/sanitize_run.py   (one tool per GPU call)
    compute-sanitizer --tool synccheck  python tools_dev/sanitize_run.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spef_b200.engine import Engine
from spef_b200.tools import synthetic
from spef_b200.spe.classification_utils import OrientationSoftClassification, PositionSoftClassification
import numpy as np

B = 4
eng = Engine(240, 384, 1728, 3, False, "bf16", B, None, 0)
eng.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
eng.set_ori_histogram(OrientationSoftClassification(12, 3, False).histogram)
x = synthetic.synthetic_images(B).cuda()
tg = synthetic.synthetic_targets(B, 2024)
eng.eval_reset()
eng.eval_batch(x, torch.from_numpy(tg["ori"]).cuda(), torch.from_numpy(tg["pos"]).cuda())
s = eng.eval_read()
print("eval sums", s[:4], "fused blocks", sum(1 for i in range(eng.num_blocks()) if eng.block_info(i)["fused"]), "launches", eng.launch_count())
eng.close()
if os.environ.get("SAN_TEMPORAL", "1") == "1":
    e2 = Engine(240, 384, 1728, 1000, True, "bf16", 1, None, 0)
    e2.load_state_dict(synthetic.synthetic_state_dict(1728, 1000))
    e2.set_ori_histogram(OrientationSoftClassification(12, 3, False).histogram)
    e2.set_pos_histogram(PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40])).histogram)
    e2.temporal_reset(1)
    os.environ["SPEF_TEMPORAL_GRAPH"] = "0"
    for _ in range(2):
        t = e2.temporal_step(x[:1])
    print("temporal quat", t["video_quat"].cpu().numpy())
torch.cuda.synchronize()
print("ok")
