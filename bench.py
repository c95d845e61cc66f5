#!/usr/bin/env python
"""bench.py -- SPEF pose-inference hot path on B200: images/sec, roofline fraction, CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port), rank 0 only

A "step" is one pass of the hot path over one batch of synthetic SPEED-shaped images: Mobile-URSONet forward
(BF16 activations, tcgen05 pointwise GEMMs) -> softmax -> soft-classification decode -> pose error / ESA sums.
Workload at every N: BASELINE.json configs[1]/[2] -- batch 256 per GPU (weak scaling: images are independent, no
data-path collective; one NCCL all-reduce of the 8 float64 ESA accumulators at the end of the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Anything a library prints to fd 1 (e.g. the NCCL version banner) goes to stderr; the JSON line uses the saved descriptor.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    def np_default(o):
        if isinstance(o, np.generic):
            return o.item()
        raise TypeError(f"not JSON serializable: {type(o).__name__}")
    _REAL_STDOUT.write(json.dumps(obj, default=np_default) + "\n")
    _REAL_STDOUT.flush()


METRIC = "images/sec (device-timed, max over ranks)"
UNIT = "images/s"
IMG = (240, 384)
N_ORI = 1728


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--lanes", type=int, default=2, help="device contexts / CUDA streams the steps are issued over round-robin (Engine.lanes); 1 = one stream")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--pw-impl", type=int, default=0, help="0 tcgen05 (default), 1 SIMT cross-check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the compact decode-sweep / temporal / fp32 / eager-baseline sub-objects")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--layers", action="store_true", help="also print the per-layer table to stderr")
    ap.add_argument("--workload", default="forward", choices=["forward", "decode", "temporal", "ingest"],
                    help="forward: BASELINE configs[1]/[2] (default, the driver's line); decode: configs[3] bins-per-axis sweep of the "
                         "decode kernel alone; temporal: configs[4] D-SPEED-shaped stream, batch-1 latency and 64-stream throughput")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.gpu)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "window": "25 ms samples over a 0.4 s pre-roll of the same workload plus the timed region"}


# ------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's PyTorch + NumPy path on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_hot_path_factory():
    """One step of the reference's hot path on the CPU (SPETorch.predict + get_score, spe_torch.py:41-76,
    spe_utils.py:104-159) as restated by oracle/spef_oracle.py -- the only place bench.py executes oracle/."""
    from oracle import spef_oracle as O
    from spef_b200.tools import synthetic
    sd = synthetic.synthetic_state_dict(N_ORI, 3)
    hist = O.ori_histogram(12)[0]

    def step(images, targets):
        ori, pos = O.forward_fp32(sd, images)
        soft = O.softmax(ori.numpy())
        q, _ = O.ori_decode_batch(soft, hist)
        return O.get_score(targets, {"ori": q, "pos": pos.numpy()})
    return step


def time_cpu(seconds, sample_batch=32):
    from spef_b200.tools import synthetic
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = cpu_hot_path_factory()
    x = synthetic.synthetic_images(sample_batch)
    tg = synthetic.synthetic_targets(sample_batch)
    step(x[:4], {k: v[:4] for k, v in tg.items()})  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        step(x, tg)
        n += sample_batch
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return n / el, cores, f"{n} images in batches of {sample_batch} ({el:.1f} s), FP32, torch {torch.__version__} CPU + NumPy, {cores} threads"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from spef_b200.tools import synthetic
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = cpu_hot_path_factory()
    sample = 32  # bounded sample of the 256-image batch per step (the full batch takes ~10 s on a host CPU)
    x = synthetic.synthetic_images(sample)
    tg = synthetic.synthetic_targets(sample)
    for _ in range(max(args.warmup, 1)):
        step(x[:8], {k: v[:8] for k, v in tg.items()})
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(x, tg)
    el = time.perf_counter() - t0
    v = sample * args.steps / el
    desc = f"{sample}-image sample of the {args.batch}-image batch per step, FP32, {cores} host threads"
    workload = (f"Mobile-URSONet batched inference (forward + softmax + decode + ESA score) on a {sample}-image sample per step of the "
                f"batch-{args.batch}-per-GPU workload (the full batch takes ~10 s per step on host cores), 240x384, 1728 orientation bins")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample_images_per_step": sample,
                   "note": "reference = the reference's own PyTorch/NumPy CPU path as restated by oracle/ (the reference is pure Python and has no pip-installable package); rank 0 only"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------------
# host placement and the H2D roof (the end-to-end number is PCIe-bound: 283 MB of float images per step)
# ------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_gpu):
    """Pin this rank's threads (and therefore its first-touch pinned buffers) to the NUMA node of its GPU's PCIe root.
    Best effort: returns a description of what was done."""
    try:
        pr = torch.cuda.get_device_properties(local_gpu)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read().strip())
        cpus = open(base + "/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if node >= 0 and ids:
            os.sched_setaffinity(0, ids)
            return {"gpu_pci": bus, "numa_node": node, "cpus_bound": len(ids)}
        return {"gpu_pci": bus, "numa_node": node, "cpus_bound": 0, "note": "no NUMA information for this device: affinity unchanged"}
    except Exception as ex:  # containers often hide /sys/bus/pci
        return {"note": f"NUMA binding unavailable ({type(ex).__name__})"}


def h2d_roof(host_pinned, dev, dist, reps=4):
    """GB/s of a plain cudaMemcpyAsync of the step's image buffer on one stream per GPU: rank 0 alone, then all ranks at once."""
    dst = torch.empty_like(host_pinned, device=dev)
    nbytes = host_pinned.numel() * host_pinned.element_size()

    def once():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            dst.copy_(host_pinned, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9

    once()
    rank = dist.get_rank() if dist is not None else 0
    alone = None
    if dist is not None:
        dist.barrier()
    if rank == 0:
        alone = once()
    if dist is not None:
        dist.barrier()
        conc = once()
        t = torch.tensor([conc], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        conc = float(t.item())
        a = torch.tensor([alone if alone is not None else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        alone = float(a.item())
    else:
        conc = alone
    del dst
    return alone, conc


def torch_eager_forward(sd, x):
    """The reference network as PLAIN PyTorch library calls (cuDNN / cuBLAS through torch.nn.functional): what SPETorch runs on a
    GPU today (spe_torch.py:57-61) and the only "Blackwell path" the reference has.  Library baseline, not this
    repo's kernels and not the oracle: BatchNorm unfolded, ReLU separate, exactly the module graph of mobilenet_v2.py."""
    import torch.nn.functional as F
    settings = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def cba(y, prefix, stride, groups, act):
        w = sd[prefix + ".0.weight"]
        y = F.conv2d(y, w, None, stride, (w.shape[-1] - 1) // 2, 1, groups)
        y = F.batch_norm(y, sd[prefix + ".1.running_mean"], sd[prefix + ".1.running_var"], sd[prefix + ".1.weight"], sd[prefix + ".1.bias"], False, 0.0, 1e-5)
        return F.relu(y) if act else y

    x = cba(x, "features.features.0", 2, 1, True)
    cin, idx = 32, 1
    for t, c, n, s_ in settings:
        for i in range(n):
            stride, hidden, p, j, y = (s_ if i == 0 else 1), cin * t, f"features.features.{idx}.conv", 0, x
            if t != 1:
                y = cba(y, f"{p}.{j}", 1, 1, True); j += 1
            y = cba(y, f"{p}.{j}", stride, hidden, True); j += 1
            y = cba(y, f"{p}.{j}", 1, 1, False)
            x = x + y if (stride == 1 and cin == c) else y
            cin, idx = c, idx + 1
    x = cba(x, "features.features.18", 1, 1, True)
    f = x.mean([2, 3])
    return F.linear(f, sd["head.ori.1.weight"], sd["head.ori.1.bias"]), F.linear(f, sd["head.pos.0.weight"], sd["head.pos.0.bias"])


def gpu_eager_baseline(sd, images, steps=5):
    """images/s of the forward alone through PyTorch eager on the same GPU, bf16 + channels_last (and fp32 for context)."""
    out = {}
    for name, dt in (("bf16_channels_last", torch.bfloat16), ("fp32_channels_last", torch.float32)):
        try:
            w = {k: (v.to(images.device, dt) if v.is_floating_point() else v.to(images.device)) for k, v in sd.items()}
            x = images.to(dt).contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                for _ in range(2):
                    torch_eager_forward(w, x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    torch_eager_forward(w, x)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"images_per_s": images.shape[0] / (ms * 1e-3), "ms_per_step": ms}
            del w, x
        except Exception as ex:
            out[name] = {"unavailable": repr(ex)[:200]}
    torch.cuda.empty_cache()
    out["what"] = ("forward only (no decode / score) of the same network through torch.nn.functional (cuDNN / cuBLAS library kernels, BatchNorm "
                   "and ReLU unfused as in the reference module), same batch, CUDA-event timed; torch " + torch.__version__)
    return out


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def layer_table(eng, batch, ms):
    """Per-launch algorithmic bytes / flops next to the measured device time."""
    kinds = {0: "stem_conv3x3s2_kernel", 1: "pw_gemm", 2: "dwconv3x3_kernel", 3: "global_mean_kernel", 4: "pw_gemm"}
    esz = 2 if eng.precision == "bf16" else 4
    rows = []
    for i in range(eng.num_layers()):
        li = eng.layer_info(i)
        ein = li["hin"] * li["win"] * li["cin"] * batch
        eout = li["hout"] * li["wout"] * li["cout"] * batch
        k = li["kind"]
        if k == 0:
            by, fl = ein * 4 + eout * esz, 2 * 27 * eout
        elif k == 2:
            by, fl = (ein + eout) * esz, 2 * 9 * eout
        elif k == 1:
            by, fl = (ein + eout * (2 if li["residual"] else 1)) * esz, 2 * li["cin"] * eout
        elif k == 3:
            by, fl = (ein + eout) * esz, ein
        else:
            by, fl = ein * esz + eout * 4 + li["cin"] * li["cout"] * esz, 2 * li["cin"] * eout
        rows.append({"layer": i, "kernel": kinds[k], "cin": li["cin"], "cout": li["cout"], "hw": f"{li['hout']}x{li['wout']}",
                     "stride": li["stride"], "bytes": by, "flops": fl, "ms": float(ms[i]), "ein": ein, "eout": eout,
                     "residual": li["residual"]})
    # fused InvertedResidual blocks: one launch covers expand + depthwise + project; its compulsory traffic is the block
    # boundary (read x once, write y once) and its time is reported in the slot of the block's first layer
    merged, skip = [], set()
    fused_first = {}
    for b in range(eng.num_blocks()):
        bi = eng.block_info(b)
        if bi["fused"]:
            fused_first[bi["first_layer"]] = bi
    if eng.stem_fusion_active():
        # stem conv + first block as ONE launch (timed in the stem's slot): reads the image once, writes the block output once
        bi = eng.block_info(0)
        grp = rows[0:1 + bi["n_layers"]]
        skip.update(g["layer"] for g in grp)
        merged.append({"layer": 0, "kernel": "fused_block_kernel", "cin": 3, "cout": grp[-1]["cout"], "hw": grp[-1]["hw"], "stride": 2,
                       "bytes": grp[0]["ein"] * 4 + grp[-1]["eout"] * esz, "flops": sum(g["flops"] for g in grp),
                       "ms": sum(g["ms"] for g in grp), "hidden": grp[0]["cout"], "tile": f"{bi['tile_h']}x{4 * bi['tile_w']}",
                       "unfused_bytes": sum(g["bytes"] for g in grp), "stem_fused": True})
    if eng.pool_fusion_active():
        # last 1x1 conv + global average pool as ONE launch (timed in the conv's slot): reads x, writes the pooled vector
        for r in rows:
            if r["kernel"] == "global_mean_kernel" and r["layer"] > 0 and rows[r["layer"] - 1]["kernel"] == "pw_gemm":
                cv = rows[r["layer"] - 1]
                skip.add(r["layer"])
                cv.update({"kernel": "conv_pool_kernel", "unfused_bytes": cv["bytes"] + r["bytes"], "bytes": (cv["ein"] + r["eout"]) * esz,
                           "flops": cv["flops"] + r["flops"], "ms": cv["ms"] + r["ms"], "hw": r["hw"]})
    for r in rows:
        if r["layer"] in skip:
            continue
        bi = fused_first.get(r["layer"])
        if bi is None:
            merged.append(r)
            continue
        grp = rows[r["layer"]:r["layer"] + bi["n_layers"]]
        if bi["fused"] == 3:
            # expand conv as a GEMM (its own row), then depthwise + project in one kernel: reads the hidden tensor once, writes y
            # (+ reads the skip input); timed in the slot of the depthwise layer
            merged.append(r)
            dwl, pj = grp[1], grp[2]
            skip.update((dwl["layer"], pj["layer"]))
            merged.append({"layer": dwl["layer"], "kernel": "dw_project_kernel", "cin": dwl["cin"], "cout": pj["cout"], "hw": pj["hw"],
                           "stride": dwl["stride"], "bytes": (dwl["ein"] + pj["eout"] * (2 if pj["residual"] else 1)) * esz,
                           "flops": dwl["flops"] + pj["flops"], "ms": dwl["ms"] + pj["ms"], "tile": f"{bi['tile_h']}x{bi['tile_w']}",
                           "unfused_bytes": dwl["bytes"] + pj["bytes"]})
            continue
        skip.update(g["layer"] for g in grp)
        merged.append({"layer": r["layer"], "kernel": "fused_block_kernel", "cin": grp[0]["cin"], "cout": grp[-1]["cout"],
                       "hw": grp[-1]["hw"], "stride": max(g["stride"] for g in grp),
                       "bytes": (grp[0]["ein"] + grp[-1]["eout"]) * esz, "flops": sum(g["flops"] for g in grp),
                       "ms": sum(g["ms"] for g in grp), "hidden": grp[0]["cout"], "tile": f"{bi['tile_h']}x{bi['tile_w']}",
                       "unfused_bytes": sum(g["bytes"] for g in grp)})
    return merged


def fp32_numbers(B, sd, base, steps=5):
    """The FP32 engine (north star: logits within 1e-4 relative of the reference) at the same batch: images/s of the same step."""
    from spef_b200.engine import Engine
    from spef_b200.tools import synthetic
    from spef_b200.spe.classification_utils import OrientationSoftClassification
    dev = torch.device("cuda", torch.cuda.current_device())
    eng = Engine(IMG[0], IMG[1], N_ORI, 3, False, "fp32", B, dev)
    eng.load_state_dict(sd)
    eng.set_ori_histogram(OrientationSoftClassification(12, 3, False).histogram)
    x = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous().to(dev)
    tg = synthetic.synthetic_targets(B, 2024)
    qt, tt = torch.from_numpy(tg["ori"]).to(dev), torch.from_numpy(tg["pos"]).to(dev)
    eng.eval_reset()
    for _ in range(2):
        eng.eval_batch(x, qt, tt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.eval_batch(x, qt, tt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    by, fl = eng.forward_cost(B)
    pk = peaks()
    eng.close()
    del x
    torch.cuda.empty_cache()
    return {"config": f"Mobile-URSONet FP32 engine (CUDA-core FFMA GEMMs: tcgen05 has no FP32-input MMA and single-pass TF32 misses the 1e-4 gate), batch {B}, same step",
            "value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "dtype": "f32", "achieved_TFLOPs": fl / (ms * 1e-3) / 1e12,
            "roofline": {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": by / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                         "note": "per-layer algorithmic bytes with FP32 activations (101 MB per image); the pointwise layers of this path are FP32-FMA-issue-bound, not HBM-bound"},
            "cpu_baseline": "the line's cpu_baseline (the reference's FP32 path on the host cores) is the CPU arm of this configuration"}


def run_b200(args):
    from spef_b200.engine import Engine
    from spef_b200.tools import synthetic
    from spef_b200.tools.evaluation import reduce_eval_sums

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)   # before any pinned allocation: first touch places the host buffers on the GPU's node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    eng = Engine(IMG[0], IMG[1], N_ORI, 3, False, args.precision, B, dev, args.pw_impl)
    sd_main = synthetic.synthetic_state_dict(N_ORI, 3)
    eng.load_state_dict(sd_main)
    from spef_b200.spe.classification_utils import OrientationSoftClassification
    eng.set_ori_histogram(OrientationSoftClassification(12, 3, False).histogram)

    # synthetic batch: 32 distinct seeded images tiled to B (283 MB of f32 at B = 256 > 126 MB L2), per-rank seed
    base = synthetic.synthetic_images(min(B, 32), IMG, synthetic.IMAGE_SEED + rank)
    host_images = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous().pin_memory()
    tg = synthetic.synthetic_targets(B, 2024 + rank)
    qt_h, tt_h = torch.from_numpy(tg["ori"]).pin_memory(), torch.from_numpy(tg["pos"]).pin_memory()
    images = host_images.to(dev)
    qt, tt = qt_h.to(dev), tt_h.to(dev)
    stream = torch.cuda.current_stream(dev)
    # steps go round-robin over `lanes` device contexts with the same weights, one CUDA stream each (Engine.lanes, what evaluation()
    # does with the batches of a phase): every step is still one full pass over one 256-image batch, consecutive steps overlap on
    # the GPU (the partly filled last wave of one step's kernels next to the other step's kernels).  Each lane has its own batch.
    lanes = eng.lanes(max(1, args.lanes))
    lane_images = [images] + [images.clone() for _ in lanes[1:]]

    def lanes_begin(ev):
        for l in lanes:
            l.side_stream.wait_event(ev)

    def lanes_end():
        for l in lanes:
            stream.wait_stream(l.side_stream)

    def step(i):
        l = lanes[i % len(lanes)]
        with torch.cuda.stream(l.side_stream):
            l.eval_batch(lane_images[i % len(lanes)], qt, tt)

    def sync_all():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up -------------------------------------------------------------------------------------
    for l in lanes:
        l.eval_reset()
    sync_all()
    for i in range(max(args.warmup, 3) * len(lanes)):
        step(i)
    sync_all()

    # ---- timed region: K steps, inputs resident in HBM, CUDA events on the launching stream ----------
    # nvidia-smi needs ~0.1 s to come up and one step is 2.5 ms: the sampler starts under a pre-roll of the same workload
    # (extra untimed warm-up steps) so that it is already sampling when the timed region begins
    sampler = ClockSampler(local)
    sampler.start()
    t_pre = time.perf_counter()
    i_pre = 0
    while time.perf_counter() - t_pre < 0.4:
        step(i_pre)
        i_pre += 1
        torch.cuda.synchronize()
    for l in lanes:
        l.eval_reset()
    launches0 = sum(l.launch_count() for l in lanes)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    lanes_begin(e0)
    for i in range(args.steps):
        step(i)
    lanes_end()
    # the ESA accumulators are per context and linear: one vector per lane, added (and, at N > 1, all-reduced: the path's one
    # exchange step, SUM of the 8 float64 accumulators over NVLink)
    sums = torch.from_numpy(sum(l.eval_read() for l in lanes)).to(dev)
    if dist is not None:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    e1.record(stream)
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = sum(l.launch_count() for l in lanes) - launches0
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1000)
    s = sums.cpu().numpy()

    # ---- end to end: host buffers through the C ABI (spef_eval_batch_host): H2D of the pinned images and targets,
    # forward + decode + score, D2H of the per-image errors, every step ------------------------------------------
    # two pinned input batches alternate (a loader hands over a fresh buffer every step); the per-image results of every
    # step are read back into their own pinned buffer; spef_eval_submit_host overlaps the H2D of step i+1 with step i
    host_images2 = host_images.clone().pin_memory()
    L = len(lanes)
    per_out = [torch.empty((B, 2), dtype=torch.float32).pin_memory() for _ in range(2 * L)]

    e2e_sums = []     # the accumulated sums of every e2e_loop call, in call order (packed, plain, uint8): they must be identical

    def e2e_loop(bufs):
        """K pipelined host-buffer steps, round-robin over the lanes (each lane double-buffers its own H2D copies); wall clock
        around submit ... wait + read of the sums, max over ranks."""
        def submit(i):
            l = lanes[i % L]
            with torch.cuda.stream(l.side_stream):
                l.eval_submit_host(bufs[(i // L) % 2], qt_h, tt_h, per_out[i % (2 * L)])

        def drain():
            for l in lanes:
                with torch.cuda.stream(l.side_stream):
                    l.eval_wait()
        for i in range(2 * L):
            submit(i)
        drain()
        for l in lanes:
            with torch.cuda.stream(l.side_stream):
                l.eval_reset()
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            submit(i)
        drain()
        sums = sum(l.eval_read() for l in lanes)     # float64 accumulators of the K timed steps (the read is part of the timed region)
        sync_all()
        el = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e_sums.append(sums)
        return el

    pack_on, pack_threads = eng.host_pack_info()
    e2e_s = e2e_loop([host_images, host_images2])
    # the same loop with the plain float copy (spef_set_host_pack(0)): what the packed upload is measured against
    e2e_plain = None
    if pack_on:
        for l in lanes:
            l.set_host_pack(False)
        plain_s = e2e_loop([host_images, host_images2])
        for l in lanes:
            l.set_host_pack(True)
        e2e_plain = {"value": world * B * args.steps / plain_s, "unit": UNIT, "h2d_bytes_per_step": int(host_images.numel() * 4 + B * 28),
                     "d2h_bytes_per_step": int(B * 8), "note": "spef_set_host_pack(0): the float images cross the bus as they are",
                     "sums_equal_packed": bool(np.array_equal(e2e_sums[0], e2e_sums[1]))}
    # same loop with uint8 pixels (the input side of the path: ToTensor's /255 moves into the stem; 4x fewer H2D bytes)
    e2e_u8 = None
    if args.precision == "bf16" and args.pw_impl == 0:
        u8a = (host_images * 255).round().to(torch.uint8).pin_memory()
        u8b = u8a.clone().pin_memory()
        for l in lanes:
            l.set_image_dtype(torch.uint8)
        u8_s = e2e_loop([u8a, u8b])
        for l in lanes:
            l.set_image_dtype(torch.float32)
        e2e_u8 = {"value": world * B * args.steps / u8_s, "unit": UNIT, "h2d_bytes_per_step": int(u8a.numel() + B * 28),
                  "d2h_bytes_per_step": int(B * 8), "note": "uint8 host images (spef_set_image_dtype(SPEF_IMG_U8)); results bit-identical to float images of the same pixels"}
    # the H2D roof this end-to-end number runs against: a plain cudaMemcpyAsync of the same pinned 283 MB buffer, rank 0 alone and
    # all ranks at once (on a shared host the concurrent figure is what caps N-GPU end-to-end scaling, not the kernels)
    roof_alone, roof_conc = h2d_roof(host_images, dev, dist)
    pack_f = eng.host_pack_stats()["packed_fraction"] if pack_on else 0.0     # of the last submit (the split follows the measured rates)
    h2d_bytes = int(host_images.numel() * (4 - 2 * pack_f) + B * 28)
    e2e_val = world * B * args.steps / e2e_s
    e2e = {"value": e2e_val, "unit": UNIT,
           "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(B * 8),
           "h2d_roof_GBps": roof_conc, "h2d_roof_alone_GBps": roof_alone,
           "achieved_h2d_GBps_per_gpu": e2e_val / world / B * h2d_bytes / 1e9,
           "frac_of_h2d_roof": (e2e_val / world / B * h2d_bytes / 1e9) / roof_conc if roof_conc else None,
           "host_pack": ({"threads": pack_threads, "measured": eng.host_pack_stats(), "host_bytes_read_per_step": int(host_images.numel() * 4),
                          "what": "the caller's float images (1.11 MB each) are rounded to BF16 by the library's host threads -- the stem's own first "
                                  "step, bit-identical results -- and cross PCIe at half the bytes, chunk by chunk behind the conversion; the rest of the batch "
                                  "crosses as float meanwhile; the split balances the measured conversion and copy rates"}
                         if pack_on else None),
           "bound": ("host side: the conversion threads and the PCIe copy share the host's memory bandwidth" if pack_on else
                     "host-to-device copy (PCIe): float images are 1.11 MB each") + "; the device-timed `value` is what the kernels sustain",
           "numa": numa,
           "api": "spef_eval_submit_host / spef_eval_wait (pinned host images + targets in, per-image errors out every step; H2D of step i+1 overlaps the kernels of step i)"}

    # ---- per-kernel roofline: per-layer CUDA events on the same stream, same inputs, K steps -------------------
    nl = eng.num_layers()
    reps = max(5, min(args.steps, 11))
    ms_layers = np.median(np.stack([eng.forward_timed(images)[2] for _ in range(reps)]), axis=0).astype(np.float64)   # median: one event glitch cannot end up in the table
    pk = peaks()
    rows = layer_table(eng, B, ms_layers)
    fam = {}
    for r in rows:
        name = r["kernel"] if r["kernel"] != "pw_gemm" else ("pw_gemm_tcgen05_kernel" if (args.precision == "bf16" and args.pw_impl == 0) else "pw_gemm_simt_kernel")
        f = fam.setdefault(name, {"launches": 0, "bytes": 0.0, "flops": 0.0, "ms": 0.0})
        f["launches"] += 1; f["bytes"] += r["bytes"]; f["flops"] += r["flops"]; f["ms"] += r["ms"]
    kernels = []
    for name, f in fam.items():
        gbs = f["bytes"] / (f["ms"] * 1e-3) / 1e9 if f["ms"] > 0 else 0.0
        kernels.append({"kernel": name, "launches_per_step": f["launches"], "ms_per_step": f["ms"], "share": f["ms"] / ms_layers.sum(),
                        "achieved_GBps": gbs, "hbm_frac": gbs / pk["hbm_gbs"], "achieved_TFLOPs": f["flops"] / (f["ms"] * 1e-3) / 1e12 if f["ms"] > 0 else 0.0})
    kernels.sort(key=lambda k: -k["ms_per_step"])
    # the dominant kernel = the single launch with the largest share of the step (one shape, so that one ncu capture describes it)
    top = max(rows, key=lambda r: r["ms"])
    top_name = top["kernel"] if top["kernel"] != "pw_gemm" else "pw_gemm_tcgen05_kernel"
    top_gbs = top["bytes"] / (top["ms"] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:   # DRAM bytes of that launch from the committed `ncu --set full` capture (profiles/ncu_traffic.json), if there is one
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            ent = json.load(f).get(f"L{top['layer']:02d}") if B == 256 else None
        if ent:
            traffic, traffic_src = ent["dram_bytes"], ent["source"]
    except OSError:
        pass
    fused = top["kernel"] == "fused_block_kernel"
    roofline = {"kernel": top_name, "launch": f"layer {top['layer']}: {top['cin']}->{top.get('hidden', '')}->{top['cout']} @{top['hw']} s{top['stride']}" if fused
                else f"layer {top['layer']}: {top['cin']}->{top['cout']} @{top['hw']} s{top['stride']}",
                "bound": "hbm", "achieved": top_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": top_gbs / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": top["bytes"], "avg_launch_ms": top["ms"],
                "share_of_step": top["ms"] / ms_layers.sum(),
                "achieved_TFLOPs": top["flops"] / (top["ms"] * 1e-3) / 1e12,
                "how": f"algorithmic bytes (DESIGN.md section 5: a fused block reads x once and writes y once) / CUDA-event time of that launch on the "
                       f"launch stream, median of {reps} passes after the timed region, same inputs; share_of_step and kernels[].share are relative to "
                       f"the sum of the per-launch event times (a few % above ms_per_step: an event pair per launch adds gaps)",
                "note": ("fused InvertedResidual kernel: the 6x hidden tensor never reaches HBM, so the launch is far below the HBM roof by design; "
                         "its limiter is instruction issue / per-warp latency of the depthwise worker warps and the hand-offs between the single-warp roles "
                         "(profiles/r02_*, DESIGN.md section 4.2); unfused_bytes is what the three per-layer kernels would move") if fused else None,
                "unfused_bytes_per_launch": top.get("unfused_bytes")}
    tot_bytes, tot_flops = eng.forward_cost(B)
    shipped_bytes = float(sum(r["bytes"] for r in rows))
    step_ms = ms_total / args.steps

    # ---- multi-GPU parity on hardware: ONE fixed seed-defined dataset, sharded over the ranks + one NCCL SUM
    # all-reduce of the float64 accumulators, against rank 0 evaluating all of it alone.  Gate: rel <= 1e-12 on the sums.
    parity = None
    if dist is not None:
        n_batches, pb = 8, B
        def dataset_batch(k):
            return (synthetic.synthetic_images(min(pb, 32), IMG, 4242 + k).repeat((pb + 31) // 32, 1, 1, 1)[:pb].contiguous().to(dev),
                    synthetic.synthetic_targets(pb, 777 + k))
        eng.eval_reset()
        for k in range(rank, n_batches, world):
            xk, tk = dataset_batch(k)
            eng.eval_batch(xk, torch.from_numpy(tk["ori"]).to(dev), torch.from_numpy(tk["pos"]).to(dev))
        sharded = torch.from_numpy(eng.eval_read()).to(dev)
        dist.all_reduce(sharded, op=dist.ReduceOp.SUM)
        sharded = sharded.cpu().numpy()
        if rank == 0:
            eng.eval_reset()
            for k in range(n_batches):
                xk, tk = dataset_batch(k)
                eng.eval_batch(xk, torch.from_numpy(tk["ori"]).to(dev), torch.from_numpy(tk["pos"]).to(dev))
            alone = eng.eval_read()
            rel = float(np.max(np.abs(sharded[:4] - alone[:4]) / np.maximum(np.abs(alone[:4]), 1e-300)))
            parity = {"images": int(alone[3]), "rel_diff": rel, "gate": 1e-12, "ok": bool(rel <= 1e-12 and sharded[3] == alone[3]),
                      "esa_sharded": float((sharded[0] + sharded[1]) / sharded[3]), "esa_one_rank": float((alone[0] + alone[1]) / alone[3]),
                      "what": f"{n_batches} seed-defined batches of {pb}: ranks evaluate batches r, r+N, ... and all-reduce (NCCL SUM) the 8 float64 sums; "
                              "rank 0 then evaluates all batches alone; max relative difference of the four sums"}
        dist.barrier()

    # ---- compact secondary workloads on the driver's line (N = 1 only; each bounded to a few seconds) ------------------------
    extras = {}
    if world == 1 and not args.no_extras:
        del images
        torch.cuda.empty_cache()
        cpu = not args.no_cpu_baseline
        try:
            extras["decode_sweep"] = {"config": "BASELINE configs[3]: decode kernel isolated, bins per axis 8..32, ~1 GB of logits per size (8x the L2; batch per row)",
                                      "rows": decode_sweep(10, 1 << 30, cpu_images=32 if cpu else 0)}
        except Exception as ex:
            extras["decode_sweep"] = {"error": repr(ex)[:300]}
        try:
            t = temporal_numbers(100, 20, cpu_frames=10 if cpu else 0)
            t["config"] = "BASELINE configs[4]: temporal stream, Mobile-URSONet+ heads, batch-1 latency and 64 streams, through Engine.temporal_step"
            extras["temporal"] = t
        except Exception as ex:
            extras["temporal"] = {"error": repr(ex)[:300]}
        try:
            extras["fp32"] = fp32_numbers(B, sd_main, base)
        except Exception as ex:
            extras["fp32"] = {"error": repr(ex)[:300]}
        try:
            xb = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:B].contiguous().to(dev)
            extras["gpu_eager_baseline"] = gpu_eager_baseline(sd_main, xb)
            del xb
        except Exception as ex:
            extras["gpu_eager_baseline"] = {"error": repr(ex)[:300]}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"Mobile-URSONet {args.precision.upper()} batched inference (forward + softmax + soft-classification decode + ESA score), "
                               f"batch {B} per GPU, 3x240x384 f32 NCHW images, 1728 orientation bins, regression position head (BASELINE.json configs[1]"
                               + ("/[2]" if world > 1 else "") + ")",
                   "weights": "calibrated random init (seed 7)", "l2": "inputs larger than L2 (283 MB image batch, GB-scale activations)",
                   "parallelism": f"batch-sharded x{world}, one NCCL all-reduce of 8 f64 at the end" if world > 1 else "single GPU",
                   "pw_impl": "tcgen05" if (args.precision == "bf16" and args.pw_impl == 0) else "simt",
                   "lanes": len(lanes),
                   "lanes_note": "the K timed steps (each one full pass over one batch, each replayed as one CUDA graph by the library) are issued "
                                 "round-robin over this many device contexts / CUDA streams (Engine.lanes, as evaluation() issues the batches of a "
                                 "phase), so consecutive steps overlap on the GPU; ms_per_step = timed region / K; --lanes 1 is one stream"},
        "e2e": e2e, "e2e_plain_copy": e2e_plain, "e2e_uint8_input": e2e_u8, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "step_roofline": {"algorithmic_bytes_per_step": shipped_bytes, "flops_per_step": tot_flops,
                          "achieved_GBps": shipped_bytes / (step_ms * 1e-3) / 1e9, "hbm_frac": shipped_bytes / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                          "achieved_TFLOPs": tot_flops / (step_ms * 1e-3) / 1e12,
                          "per_layer_bytes_per_step": tot_bytes,
                          "per_layer_equivalent_GBps": tot_bytes / (step_ms * 1e-3) / 1e9,
                          "note": "algorithmic bytes of the decomposition that ships (fused blocks move only their boundary tensors); "
                                  "per_layer_* is the traffic of the reference's layer-by-layer decomposition (SURVEY 8d: 51.19 MB/image) "
                                  "over the same time, i.e. the HBM rate a per-layer implementation would need to match this step"},
        "kernels": kernels,
        "esa": {"images": float(s[3]), "esa_score": float((s[0] + s[1]) / s[3]), "flagged": float(s[4] + s[5])},
    }
    if parity is not None:
        out["multi_gpu_parity"] = parity
    out.update(extras)
    if rank == 0 and not args.no_cpu_baseline and world >= 1:
        v, cores, desc = time_cpu(args.cpu_seconds)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
    if rank == 0:
        if args.layers:
            for r in rows:
                gb = r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] > 0 else 0
                print(f"L{r['layer']:02d} {r['kernel']:24s} {r['cin']:5d}->{r['cout']:5d} {r['hw']:>8s} s{r['stride']} {r['ms']*1000:9.1f} us "
                      f"{gb:8.0f} GB/s {r['flops']/(r['ms']*1e-3)/1e12 if r['ms'] > 0 else 0:7.1f} TF/s", file=sys.stderr)
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def decode_sweep(steps, total_bytes=1 << 29, cpu_images=0):
    """BASELINE configs[3]: soft-classification decode kernel isolated, orientation bins per axis 8..32.  Returns one row per
    size: images/s and achieved GB/s of the algorithmic bytes (4 n_bins + 16 per image) against the measured HBM roof;
    cpu_images > 0 adds the oracle's decode_batch (the reference's per-image NumPy + LAPACK loop) on that many images."""
    from spef_b200.engine import Engine
    from spef_b200.spe.classification_utils import OrientationSoftClassification
    from spef_b200._ffi import ptr
    dev = torch.device("cuda", torch.cuda.current_device())
    eng = Engine(32, 32, 8, 3, False, "fp32", 1, dev)
    pk = peaks()
    rows = []
    for n_dim in (8, 12, 16, 24, 32):
        hist = OrientationSoftClassification(n_dim, 3, False).histogram
        n = hist.shape[0]
        eng.set_ori_histogram(hist)
        # logits larger than L2, in whole multiples of the persistent grid's warps (148 CTAs x 16 warps, one image per warp and
        # pass) and, from 32 passes up, of the 32-image eigen-solve batches
        wave = 16 * torch.cuda.get_device_properties(dev).multi_processor_count
        passes = max(1, round(total_bytes / (4 * n) / wave))
        if passes >= 32:
            passes -= passes % 32
        B = int(passes * wave)
        logits = torch.randn((B, n), device=dev) * 3
        # pre-allocated outputs and the bare C-ABI call in the timed loop: the Python wrapper's allocations would make the
        # GPU wait for the host at these kernel durations
        quat = torch.empty((B, 4), device=dev)
        flags = torch.zeros(B, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream or None

        def call():
            rc = eng.lib.spef_decode_ori(eng._h, ptr(logits), B, n, 1, None, ptr(quat), None, None, ptr(flags), st)
            assert rc == 0
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = B * (4 * n + 16) / (ms * 1e-3) / 1e9
        row = {"bins_per_axis": n_dim, "n_bins": n, "batch": B, "ms": ms, "images_per_s": B / (ms * 1e-3),
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None}}
        if cpu_images > 0:  # the reference's decode on the host cores (oracle restatement; cpu_baseline leg)
            from oracle import spef_oracle as O
            z = logits[:cpu_images].cpu().numpy()
            t0 = time.perf_counter()
            O.ori_decode_batch(O.softmax(z), np.asarray(hist, np.float64))
            el = time.perf_counter() - t0
            row["cpu_baseline"] = {"value": cpu_images / el, "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": f"{cpu_images} images, softmax + decode_batch (NumPy + LAPACK eig per image)"}
        rows.append(row)
        del logits
    eng.close()
    return rows


def run_decode_sweep(args):
    rows = decode_sweep(args.steps, 1 << 29, cpu_images=0 if args.no_cpu_baseline else 64)
    emit({"metric": "decode images/sec (softmax + weighted quaternion average, kernel isolated)", "unit": UNIT, "n_gpus": 1,
          "steps": args.steps, "warmup": 3, "dtype": "f32", "data": "synthetic", "higher_is_better": True,
          "value": rows[1]["images_per_s"], "config": {"workload": "BASELINE configs[3]: decode sweep, bins per axis 8..32, logits ~ 3*N(0,1), inputs larger than L2"},
          "algorithmic_bytes": "4*n_bins + 16 per image", "peak_GBps": peaks()["hbm_gbs"], "sweep": rows})


def temporal_numbers(n_frames_b1=200, n_steps_s64=50, cpu_frames=0):
    """BASELINE configs[4]: Mobile-URSONet+ (1728 + 1000 bins) frame stream through spef_temporal_step (forward + softmax + decode +
    adaptive pdf filter + decode): batch-1 latency per frame and 64 parallel streams throughput, both through the call a user makes
    (Engine.temporal_step: the library replays the few-stream step as ONE CUDA graph after the second frame)."""
    from spef_b200.engine import Engine
    from spef_b200.tools import synthetic
    from spef_b200.spe.classification_utils import OrientationSoftClassification, PositionSoftClassification
    dev = torch.device("cuda", torch.cuda.current_device())
    sd = synthetic.synthetic_state_dict(N_ORI, 1000)
    ori_hist = OrientationSoftClassification(12, 3, False).histogram
    pos_hist = PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40])).histogram
    pk = peaks()
    out = {}
    for S in (1, 64):
        eng = Engine(IMG[0], IMG[1], N_ORI, 1000, True, "bf16", S, dev)
        eng.load_state_dict(sd)
        eng.set_ori_histogram(ori_hist)
        eng.set_pos_histogram(pos_hist)
        eng.temporal_reset(S)
        frames = synthetic.synthetic_images(S, IMG, 99).to(dev)
        for _ in range(5):
            eng.temporal_step(frames)
        torch.cuda.synchronize()
        lat = []
        for _ in range(n_frames_b1 if S == 1 else n_steps_s64):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.temporal_step(frames)
            e1.record()
            e1.synchronize()
            lat.append(e0.elapsed_time(e1))
        lat = np.array(lat)
        key = "batch1" if S == 1 else "streams64"
        l0 = eng.launch_count()
        eng.temporal_step(frames)
        by, fl = eng.forward_cost(S)
        by += 8.9e6                                     # + the BF16 weights, read once per step (not amortised at batch 1)
        gbs = by / (np.median(lat) * 1e-3) / 1e9
        out[key] = {"p50_ms": float(np.percentile(lat, 50)), "p99_ms": float(np.percentile(lat, 99)),
                    "frames_per_s": float(S / (np.median(lat) * 1e-3)), "kernels_per_frame_step": eng.launch_count() - l0,
                    "cuda_graph_inside_the_library": bool(S <= int(os.environ.get("SPEF_TEMPORAL_GRAPH_MAX", "256")) and os.environ.get("SPEF_TEMPORAL_GRAPH", "1") != "0"),
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"], "traffic": None,
                                 "note": "per-layer algorithmic bytes of one forward at this batch + the weights; batch 1 is latency-bound (one frame cannot fill 148 SMs)"}}
        eng.close()
    if cpu_frames > 0:  # the reference's frame loop on the host cores (oracle restatement; cpu_baseline leg)
        from oracle import spef_oracle as O
        cores = len(os.sched_getaffinity(0))
        torch.set_num_threads(cores)
        ref = O.TemporalInference(O.ori_histogram(12)[0], O.pos_histogram(10))
        x = synthetic.synthetic_images(1, IMG, 99)
        O.forward_fp32(sd, x)
        t0 = time.perf_counter()
        for _ in range(cpu_frames):
            o, p_ = O.forward_fp32(sd, x)
            ref.step(o[0].numpy(), p_[0].numpy())
        el = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cpu_frames / el, "unit": "frames/s", "cores": cores, "kind": "port",
                               "sample": f"{cpu_frames} frames, batch 1: FP32 forward + softmax + decode + TemporalPDF filters + decode ({el / cpu_frames * 1e3:.1f} ms per frame)"}
    return out


def run_temporal(args):
    out = {"metric": "frames/sec and per-frame latency (forward + softmax + decode + adaptive pdf filter + decode)", "unit": "frames/s",
           "n_gpus": 1, "dtype": "bf16", "data": "synthetic", "higher_is_better": True,
           "config": {"workload": "BASELINE configs[4]: temporal D-SPEED-shaped stream, Mobile-URSONet+ heads, random frames"}}
    out.update(temporal_numbers(cpu_frames=0 if args.no_cpu_baseline else 20))
    out["value"] = out["streams64"]["frames_per_s"]
    emit(out)


def run_ingest(args):
    """Input side of the path: decoded 1200 x 1920 greyscale camera frames -> Resize((240, 384)) -> ToTensor
    (spef_resize_frames, bit-exact against torchvision + Pillow) -> forward + decode + score.  Reports the resize kernel alone
    (device-resident frames, against the HBM roof), the frames -> score step on the device, the same step end to end from
    pinned host frames, and the reference's CPU transform (PIL resize + ToTensor restated in the oracle) on one host core."""
    from spef_b200.engine import Engine
    from spef_b200.tools import synthetic
    from spef_b200.spe.classification_utils import OrientationSoftClassification
    dev = torch.device("cuda", 0)
    B, SH, SW = args.batch, 1200, 1920
    eng = Engine(IMG[0], IMG[1], N_ORI, 3, False, "bf16", B, dev)
    eng.load_state_dict(synthetic.synthetic_state_dict(N_ORI, 3))
    eng.set_ori_histogram(OrientationSoftClassification(12, 3, False).histogram)
    eng.set_image_dtype(torch.uint8)
    base = torch.from_numpy(synthetic.synthetic_frames(8, SH, SW, 1, seed=21, kind="speed"))
    host = base.repeat((B + 7) // 8, 1, 1)[:B].contiguous().pin_memory()          # 590 MB at B = 256
    frames = host.to(dev)
    tg = synthetic.synthetic_targets(B, 2024)
    qt, tt = torch.from_numpy(tg["ori"]).to(dev), torch.from_numpy(tg["pos"]).to(dev)
    pk = peaks()

    def timed(fn, steps):
        for _ in range(max(3, args.warmup)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    out_u8 = torch.empty((B, 3, IMG[0], IMG[1]), dtype=torch.uint8, device=dev)
    from spef_b200._ffi import ptr
    st = torch.cuda.current_stream(dev).cuda_stream or None

    def resize_only():
        assert eng.lib.spef_resize_frames(eng._h, ptr(frames), B, SH, SW, 1, ptr(out_u8), 1, st) == 0

    ms_rz = timed(resize_only, args.steps)
    alg = B * (SH * SW + 3 * IMG[0] * IMG[1])
    eng.eval_reset()

    def frames_to_score():
        resize_only()
        eng.eval_batch(out_u8, qt, tt)

    ms_dev = timed(frames_to_score, args.steps)

    # end to end from pinned host frames: the H2D copy of step i + 1 (590 MB, PCIe) runs on a copy stream while step i computes
    # (two device frame buffers, events in both directions)
    dbuf = [frames, torch.empty_like(frames)]
    copy_stream = torch.cuda.Stream(dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    main = torch.cuda.current_stream(dev)

    def submit_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dbuf[i % 2].copy_(host, non_blocking=True)
            copied[i % 2].record(copy_stream)

    def e2e_run(n):
        for ev in consumed:
            ev.record(main)
        submit_copy(0)
        for i in range(n):
            if i + 1 < n:
                submit_copy(i + 1)
            main.wait_event(copied[i % 2])
            assert eng.lib.spef_resize_frames(eng._h, ptr(dbuf[i % 2]), B, SH, SW, 1, ptr(out_u8), 1, st) == 0
            consumed[i % 2].record(main)
            eng.eval_batch(out_u8, qt, tt)

    e2e_run(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    sums = eng.eval_read()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # CPU: the reference's transform (PIL antialiased bilinear + ToTensor) on one core, a bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from PIL import Image
            from torchvision import transforms
            tf = transforms.Compose([transforms.Resize(IMG), transforms.ToTensor()])
            fr = [Image.fromarray(f).convert("RGB") for f in base.numpy()]
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < min(args.cpu_seconds, 5.0):
                tf(fr[n % len(fr)])
                n += 1
            cpu = {"value": n / (time.perf_counter() - t0), "unit": "frames/s", "cores": 1, "kind": "reference",
                   "sample": f"{n} frames through torchvision Resize((240, 384)) + ToTensor on PIL RGB images (the reference's SPEDataset transform; JPEG decode excluded)"}
        except Exception as ex:  # torchvision / Pillow missing on the box
            cpu = {"unavailable": repr(ex)}
    emit({"metric": "images/sec (camera frames -> resize -> forward + decode + score)", "unit": UNIT, "n_gpus": 1, "steps": args.steps,
          "warmup": max(3, args.warmup), "dtype": "u8 resize (22-bit fixed point) + bf16 network", "data": "synthetic", "higher_is_better": True,
          "value": B / (ms_dev * 1e-3), "ms_per_step": ms_dev,
          "config": {"workload": f"1200x1920 greyscale uint8 frames, batch {B}: spef_resize_frames -> uint8 [B,3,240,384] -> spef_eval_batch",
                     "l2": "inputs larger than L2 (590 MB of frames per step)"},
          "e2e": {"value": B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(host.numel()), "d2h_bytes_per_step": 64,
                  "api": "pinned host frames -> H2D on a copy stream (double buffered, overlapping the previous step) -> spef_resize_frames -> spef_eval_batch, sums read back at the end"},
          "roofline": {"kernel": "resize_aa_kernel", "bound": "hbm", "achieved": alg / (ms_rz * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                       "frac": alg / (ms_rz * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "algorithmic_bytes_per_launch": alg,
                       "avg_launch_ms": ms_rz, "peak_source": pk["source"], "frames_per_s": B / (ms_rz * 1e-3)},
          "cpu_baseline": cpu, "esa": {"images": float(sums[3]), "esa_score": float((sums[0] + sums[1]) / max(sums[3], 1.0))}})


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "decode":
        run_decode_sweep(args)
    elif args.workload == "temporal":
        run_temporal(args)
    elif args.workload == "ingest":
        run_ingest(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
