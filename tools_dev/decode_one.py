"""Developer tool: a few launches of the decode kernel at one sweep point (for ncu).  Usage: decode_one.py n_dim B [n_launches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spef_b200._ffi import ptr
from spef_b200.engine import Engine
from spef_b200.spe.classification_utils import OrientationSoftClassification
n_dim, B = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
e = Engine(32, 32, 8, 3, False, "fp32", 1)
hist = OrientationSoftClassification(n_dim, 3, False).histogram
e.set_ori_histogram(hist)
n = hist.shape[0]
logits = torch.randn((B, n), device=dev) * 3
quat = torch.empty((B, 4), device=dev)
flags = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(reps):
    assert e.lib.spef_decode_ori(e._h, ptr(logits), B, n, 1, None, ptr(quat), None, None, ptr(flags), None) == 0
torch.cuda.synchronize()
print("ok", float(quat.abs().sum()))
