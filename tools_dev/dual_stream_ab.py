"""Developer tool: K evaluation steps (B = 256 each) through ONE engine on one stream vs alternating between TWO engine contexts
on two streams (independent batches: the tail of one step's kernels overlaps the other's)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spef_b200.engine import Engine
from spef_b200.tools import synthetic
from spef_b200._ffi import ptr
from oracle import spef_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NE = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = synthetic.synthetic_state_dict(1728, 3)
hist = O.ori_histogram(12)[0]
engs, xs, streams = [], [], []
tg = synthetic.synthetic_targets(B)
qt, tt = torch.as_tensor(tg["ori"]).float().cuda(), torch.as_tensor(tg["pos"]).float().cuda()
for i in range(NE):
    e = Engine(240, 384, 1728, 3, False, "bf16", B, "cuda:0")
    e.load_state_dict(sd); e.set_ori_histogram(hist)
    engs.append(e); xs.append(synthetic.synthetic_images(B).cuda()); streams.append(torch.cuda.Stream())
torch.cuda.synchronize()
def run(n_eng, K=40):
    def go(k):
        for i in range(k):
            j = i % n_eng
            e = engs[j]
            assert e.lib.spef_eval_batch(e._h, ptr(xs[j]), ptr(qt), ptr(tt), B, None, streams[j].cuda_stream) == 0
    go(6); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams[:n_eng]: s.wait_event(e0)
    go(K)
    for s in streams[:n_eng]: main.wait_stream(s)
    e1.record(main); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for n in range(1, NE + 1):
    ms = run(n)
    print(f"{n} engine(s): {ms:.4f} ms per step, {B / ms * 1e3:.0f} img/s", flush=True)
ms = run(1); print(f"1 engine(s): {ms:.4f} ms per step, {B / ms * 1e3:.0f} img/s")
