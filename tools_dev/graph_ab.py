"""Developer tool: the B = 256 evaluation step as direct launches vs replayed as one CUDA graph (captured here with torch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spef_b200.engine import Engine
from spef_b200.tools import synthetic
from oracle import spef_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = synthetic.synthetic_state_dict(1728, 3)
eng = Engine(240, 384, 1728, 3, False, "bf16", B, "cuda:0")
eng.load_state_dict(sd)
eng.set_ori_histogram(O.ori_histogram(12)[0])
x = synthetic.synthetic_images(B).cuda()
tg = synthetic.synthetic_targets(B)
qt, tt = torch.as_tensor(tg["ori"]).float().cuda(), torch.as_tensor(tg["pos"]).float().cuda()
def step():
    eng.eval_batch(x, qt, tt)
def timeit(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("direct  ms/step", timeit(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step(); torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        step()
print("graph   ms/step", timeit(g.replay))
print("direct  ms/step", timeit(step))
