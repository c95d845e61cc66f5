"""Developer tool (GPU box): run the stride-2 depthwise -> project block (block index 13) teacher-forced at several batch sizes and
compare with the per-layer chain."""
import sys
import torch
sys.path.insert(0, ".")
from spef_b200.engine import Engine
from spef_b200.tools import synthetic

sd = synthetic.synthetic_state_dict(1728, 3)
eng = Engine(240, 384, 1728, 3, False, "bf16", 256, "cuda:0")
eng.load_state_dict(sd)
bi = 13
info = eng.block_info(bi)
print(info, flush=True)
for B in (256,) * 12:
    x = (torch.randn(B, 15, 24, 96, device="cuda") * 0.5).to(torch.bfloat16)
    got = eng.block_forward(bi, x)
    torch.cuda.synchronize()
    cur = x
    for li in range(info["first_layer"], info["first_layer"] + info["n_layers"]):
        cur = eng.layer_forward(li, cur, None)
    torch.cuda.synchronize()
    d = (got.float() - cur.float()).abs().amax(dim=(1, 2, 3))
    bad = torch.nonzero(d > 0).flatten().tolist()
    print(B, bool(torch.equal(got, cur)), float(d.max()), "bad images:", bad[:40], len(bad), flush=True)
    if bad:
        dd = (got.float() - cur.float()).abs()[bad[0]]
        print("  image", bad[0], "bad rows:", torch.nonzero(dd.amax(dim=(1, 2)) > 0).flatten().tolist(), "bad cols:", torch.nonzero(dd.amax(dim=(0, 2)) > 0).flatten().tolist(), "bad channels:", len(torch.nonzero(dd.amax(dim=(0, 1)) > 0)), flush=True)

print("forward at 256", flush=True)
x = synthetic.synthetic_images(8).repeat(32, 1, 1, 1).contiguous().cuda()
for it in range(12):
    o, p = eng.forward(x)
    torch.cuda.synchronize()
    print(it, float(o.abs().max()), flush=True)
