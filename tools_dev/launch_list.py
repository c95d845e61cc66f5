"""Developer tool (CPU box): the launches of ONE step out of an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py
(tools_dev/final_runs.sh), for profiles/.   python tools_dev/launch_list.py gpurun_out/final_launches.csv "<title>" > profiles/...txt
A step = the launches from one stem kernel (fused_block_t_kernel<..., 1, 1> or the im2col GEMM) up to the next; the last complete one is printed."""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit() and r[12] == "gpu__time_duration.sum"]
ls = []
for r in rows:
    v = float(r[14].replace(",", ""))
    us = v / 1000.0 if r[13] in ("ns", "nsecond") else (v if r[13] in ("us", "usecond") else v * 1000.0)
    ls.append((r[4], r[7], r[8], us, r[6]))
starts = [i for i, l in enumerate(ls) if "fused_block_t_kernel<1, 6, 2, 1, 1>" in l[0] or "pw_gemm_tcgen05_v2_kernel<0, 1, 4, 2>" in l[0]]
# launches of the two lanes interleave in the log: take one stream's launches between two of its stem kernels
stream = ls[starts[-2]][4]
mine = [l for l in ls[starts[0]:] if l[4] == stream]
st = [i for i, l in enumerate(mine) if "fused_block_t_kernel<1, 6, 2, 1, 1>" in l[0] or "pw_gemm_tcgen05_v2_kernel<0, 1, 4, 2>" in l[0]]
step = mine[st[-2]:st[-1]]
tot = sum(l[3] for l in step)
print(f"# {sys.argv[2]}")
print(f"# the {len(step)} launches of ONE step (B = 256) on one of the two lanes, in order; per-launch times are cold-cache and serialised (compare shares, not absolutes)")
for name, blk, grd, us, _ in step:
    print(f"{us:9.1f} us  {100 * us / tot:5.1f} %  grid {grd:>12} block {blk:>12}  {name[:150]}")
print(f"{tot:9.1f} us  total of the step's kernels under ncu")
