"""Developer tool: decode sweep timings for several logits volumes.  Usage: decode_time.py log2_bytes [log2_bytes ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
for lb in sys.argv[1:]:
    rows = bench.decode_sweep(10, 1 << int(lb), cpu_images=0)
    print(lb, " ".join(f"{r['bins_per_axis']}:{r['batch']}:{r['ms']*1e3:.1f}us:{r['roofline']['frac']:.3f}" for r in rows), flush=True)
