"""Developer diagnostic: per-image decode kernel (SPEF_DECODE_STREAM=0) vs the streaming kernel (=2) over bins per axis and
batch size, CUDA-event timed on the launch stream.  Usage: python tools_dev/decode_ab.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spef_b200._ffi import ptr  # noqa: E402
from spef_b200.engine import Engine  # noqa: E402
from spef_b200.spe.classification_utils import OrientationSoftClassification  # noqa: E402


def make(mode, cfg=1):
    os.environ["SPEF_DECODE_STREAM"] = str(mode)
    os.environ["SPEF_DECODE_CFG"] = str(cfg)
    e = Engine(32, 32, 8, 3, False, "fp32", 1)
    del os.environ["SPEF_DECODE_STREAM"], os.environ["SPEF_DECODE_CFG"]
    return e


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda", 0)
    engines = {"per_image": make(0), "s8x4": make(1, 0), "s16x2": make(1, 2)}
    st = torch.cuda.current_stream(dev).cuda_stream or None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for n_dim in (8, 12, 16, 24, 32):
        hist = OrientationSoftClassification(n_dim, 3, False).histogram
        n = hist.shape[0]
        for e in engines.values():
            e.set_ori_histogram(hist)
        bmax = int(max(4096, min(262144, (1 << 29) // (4 * n))))
        for B in sorted({256, 1024, 4096, 16384, 65536, bmax}):
            if B > bmax:
                continue
            logits = torch.randn((B, n), device=dev) * 3
            quat = torch.empty((B, 4), device=dev)
            flags = torch.zeros(B, dtype=torch.int32, device=dev)
            line = f"n_dim {n_dim:2d} n {n:6d} B {B:7d}"
            for name, e in engines.items():
                def call():
                    assert e.lib.spef_decode_ori(e._h, ptr(logits), B, n, 1, None, ptr(quat), None, None, ptr(flags), st) == 0
                for _ in range(3):
                    call()
                torch.cuda.synchronize()
                tot = 0.0
                for _ in range(steps):
                    if B * n * 4 < (200 << 20):
                        flush.zero_()   # small inputs: evict them from L2 between calls
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    call()
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                ms = tot / steps
                line += f" | {name} {ms * 1e3:8.1f} us {B * (4 * n + 16) / ms / 1e6:5.0f} GB/s"
            print(line, flush=True)
            del logits


if __name__ == "__main__":
    main()
