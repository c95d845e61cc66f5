"""Developer tool (CPU box): text summary of one `ncu --set full --import-source on` capture for profiles/.
    python tools_dev/ncu_summary.py gpurun_out/ncu_<tag>_raw.csv gpurun_out/ncu_<tag>_src.csv "<title>" > profiles/r02_ncu_<tag>.txt
Prints the headline counters, the stall-reason mix and -- for the warp-specialised kernels -- the per-role share of the issued
instructions and samples (roles are told apart by their setmaxnreg instruction) plus every mbarrier wait / TMEM / TMA / MMA site."""
import csv
import re
import sys
from collections import Counter

raw, src, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1]))
print(f"# {title}")
print(f"Kernel Name = {d.get('Kernel Name')}")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in d:
        print(f"{k} = {d[k]}")
try:
    w, c = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]), float(d["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"])
    print(f"shared-memory bank-conflict wavefronts = {100 * c / w:.1f} % of the shared wavefronts")
except Exception:
    pass
for k in sorted(d):
    if "issue_stalled" in k and "per_issue_active" in k:
        print(f"{k} = {d[k]}")

rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
names = hdr[30:47]
ins = [(r[1].strip(), int(r[4] or 0), int(r[5] or 0), [int(x or 0) for x in r[30:47]]) for r in data if len(r) > 46]
tot_s, tot_e = sum(x[1] for x in ins), sum(x[2] for x in ins)
print(f"\n# source page (SASS): {len(ins)} instructions, {tot_e} warp instructions executed, {tot_s} stall samples")
mix = Counter()
for t, s_, e, _ in ins:
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] if t else "?"
    mix[op] += e
print("dynamic instruction mix: " + "  ".join(f"{k} {100 * v / tot_e:.1f}%" for k, v in mix.most_common(16)))
b = [i for i, x in enumerate(ins) if "USETMAXREG" in x[0]]
if b:
    print("\n# per role (code between two setmaxnreg instructions; TRY_ALLOC = depthwise workers, the largest DEALLOC block = epilogue)")
    bb = [0] + b + [len(ins)]
    for k in range(len(bb) - 1):
        seg = ins[bb[k]:bb[k + 1]]
        sm, ex = sum(x[1] for x in seg), sum(x[2] for x in seg)
        st = [sum(x[3][j] for x in seg) for j in range(len(names))]
        top = sorted(zip(names, st), key=lambda t: -t[1])[:4]
        head = seg[0][0][:44] if k else "(prologue)"
        print(f"  [{bb[k]:4d},{bb[k + 1]:4d}) {head:44s} exec {ex:10d} ({100 * ex / tot_e:4.1f} %)  samples {sm:6d}  " +
              " ".join(f"{n[6:]}:{100 * v / max(sm, 1):.0f}%" for n, v in top))
print("\n# mbarrier wait sites / tensor, TMEM and TMA instructions: SASS row, warp executions, samples, instruction")
for i, (t, s_, e, _) in enumerate(ins):
    if re.search(r"SYNCS\.PHASECHK|UTCHMMA|LDTM|UTMALDG|UBLKCP|UTCBAR|NANOSLEEP", t) and e:
        print(f"{i:5d} {e:10d} {s_:6d}  {t[:90]}")
