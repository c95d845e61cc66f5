"""evaluation() (reference: src/tools/evaluation.py:35-102) with on-device score accumulation.

For a spef_b200 plug-in (SPEB200) the whole batch step -- H2D, forward, softmax, decode, pose error -- is one
C-ABI call (spef_eval_batch_host); the five ESA sums stay on the device as float64 and are read once per phase.
When torch.distributed is initialised (one process per GPU, each holding a shard of the loader) the 8 sums are
all-reduced (SUM) once per phase -- the only collective on the path -- and the per-image errors are all-gathered
for std / MAD.  Any other duck-typed back-end with predict() takes the reference's per-batch route.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import numpy as np
import torch

from ..spe.spe_b200 import SPEB200
from ..spe.spe_utils import SPEUtils
from .utils import RunningAverage


def mad(data) -> float:
    """Median absolute deviation (evaluation.py:16-32)."""
    median = np.median(data)
    return np.median(np.abs(np.array(data) - median)).tolist()


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def reduce_eval_sums(sums: np.ndarray, device=None) -> np.ndarray:
    """SUM all-reduce of the 8 float64 accumulators across ranks (NCCL on GPU tensors, gloo on CPU tensors).
    Linear, so it reproduces RunningAverage's sum_b(mean_b * n_b) / sum_b n_b over the union of the shards."""
    dist = _dist()
    if dist is None:
        return sums
    t = torch.from_numpy(np.array(sums, dtype=np.float64, copy=True))  # never reduce into the caller's array
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_per_image(err: np.ndarray, device=None) -> np.ndarray:
    """All-gather of the per-image error rows [n_local, 2] (ragged across ranks)."""
    dist = _dist()
    if dist is None:
        return err
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, err)
    return np.concatenate(parts, axis=0)


def _raise_decode_guards(sums: np.ndarray) -> None:
    """The reference's decode() raises on these conditions (classification_utils.py:134-135, 253-254, 262-263); the fused route
    counts the per-image guard flags of every batch into sums[6] / sums[7] (spef_b200.h) -- after the cross-rank reduction, so
    that every rank raises together."""
    if sums[6] > 0:
        raise ValueError("Error during orientation decoding")
    if sums[7] > 0:
        raise ValueError("Error during position decoding, NaN found in decoded position (or the encoded position vector sum is zero).")


def _finish(rec_score, rec_error, phase, sums, per_image, device=None, stats_engine=None):
    sums = reduce_eval_sums(sums, device)
    _raise_decode_guards(sums)
    per_image = gather_per_image(per_image, device)
    m = SPEUtils.metrics_from_sums(sums)
    rec_score[phase]['ori'].append(float(m['ori_score']))
    rec_score[phase]['pos'].append(float(m['pos_score']))
    rec_score[phase]['esa'].append(float(m['esa_score']))
    rec_error[phase]['ori'].append(float(m['ori_error']))
    rec_error[phase]['pos'].append(float(m['pos_error']))
    if stats_engine is not None and per_image.shape[0] > 0:
        # std / median absolute deviation on the device (spef_error_stats: float64 moments, exact radix-select medians)
        dev = torch.as_tensor(np.ascontiguousarray(per_image, np.float32)).to(stats_engine.device)
        so, sp = stats_engine.error_stats(dev, 0), stats_engine.error_stats(dev, 1)
        rec_error[phase]['ori_std'].append(so['std'])
        rec_error[phase]['pos_std'].append(sp['std'])
        rec_error[phase]['ori_mad'].append(so['mad'])
        rec_error[phase]['pos_mad'].append(sp['mad'])
        return
    rec_error[phase]['ori_std'].append(np.std(per_image[:, 0]).tolist())
    rec_error[phase]['pos_std'].append(np.std(per_image[:, 1]).tolist())
    rec_error[phase]['ori_mad'].append(mad(per_image[:, 0]))
    rec_error[phase]['pos_mad'].append(mad(per_image[:, 1]))


def evaluation(
    spe_model: Any,
    dataloader: Dict[str, Any],
    spe_utils: SPEUtils,
    split: Tuple[str, ...] = ('test', 'valid'),
    device_stats: bool = False,
) -> Tuple[Dict[str, Dict[str, List[float]]], Dict[str, Dict[str, List[float]]]]:
    """Same arguments and return structure as the reference: rec_score[phase] = {'ori','pos','esa'},
    rec_error[phase] = {'ori','pos','ori_std','pos_std','ori_mad','pos_mad'}, each a list of length 1.
    device_stats=True (SPEB200 back-end only) computes the std / MAD entries with spef_error_stats on the GPU instead of
    NumPy on the host (float64 moments instead of NumPy's float32 pairwise sums: equal to ~1e-6 relative)."""
    rec_score = {x: {'ori': [], 'pos': [], 'esa': []} for x in split}
    rec_error = {x: {'ori': [], 'pos': [], 'ori_std': [], 'pos_std': [], 'ori_mad': [], 'pos_mad': []} for x in split}

    for phase in split:
        if isinstance(spe_model, SPEB200):
            eng = spe_model.engine
            # the batches of the phase go round-robin over the lanes (device contexts with the same weights, one CUDA stream each:
            # Engine.lanes); the caller's stream only orders the loader's own device work before each batch
            lanes = eng.lanes(getattr(spe_model, "lanes", 1))
            main = torch.cuda.current_stream(eng.device)
            for l in lanes:
                l.side_stream.wait_stream(main)
                with torch.cuda.stream(l.side_stream):
                    l.eval_reset()
            per, keep = [], []
            dtype0 = eng.image_dtype
            for i, (images, targets) in enumerate(dataloader[phase]):
                l = lanes[i % len(lanes)]
                x = images['torch']
                # a loader that yields uint8 pixels (ToTensor not applied yet) takes the uint8 ingest route: a quarter of the
                # H2D bytes, bit-identical results (the stem divides by 255 itself)
                want = torch.uint8 if (x.dtype == torch.uint8 and eng.precision == "bf16") else torch.float32
                with torch.cuda.stream(l.side_stream):
                    if want != l.image_dtype:
                        l.eval_wait()
                        l.set_image_dtype(want)
                    if x.device.type == "cpu":
                        # pipelined: the H2D copy of this batch overlaps the kernels of the previous one; host buffers are kept
                        # alive until the phase is drained
                        x = l._host_img(x)
                        qt = torch.as_tensor(targets['ori']).detach().to("cpu", torch.float32).contiguous()
                        tt = torch.as_tensor(targets['pos']).detach().to("cpu", torch.float32).contiguous()
                        out = torch.empty((x.shape[0], 2), dtype=torch.float32, pin_memory=True)
                        l.eval_submit_host(x, qt, tt, out)
                        keep.append((x, qt, tt))
                        per.append(out)
                    else:
                        l.side_stream.wait_stream(main)      # the loader produced this batch on the caller's stream
                        x.record_stream(l.side_stream)
                        per.append(l.eval_batch(x, targets['ori'], targets['pos'], want_per_image=True))
                        keep.append(x)
            sums = np.zeros(8, np.float64)
            for l in lanes:
                with torch.cuda.stream(l.side_stream):
                    l.eval_wait()
                    sums += l.eval_read()
                    if l.image_dtype != dtype0:
                        l.set_image_dtype(dtype0)
                main.wait_stream(l.side_stream)
            per = [p.cpu().numpy() if isinstance(p, torch.Tensor) else p for p in per]
            del keep
            per_image = np.concatenate(per, axis=0) if per else np.zeros((0, 2), np.float32)
            _finish(rec_score, rec_error, phase, sums, per_image, eng.device, eng if device_stats else None)
        else:
            # any other back-end with predict(): per-batch route of the reference (evaluation.py:69-85); the score
            # itself still runs in libspef_b200.so through SPEUtils.get_score
            sums = np.zeros(8, np.float64)
            per = []
            running_avg = RunningAverage(keys=('esa_score', 'ori_score', 'pos_score', 'ori_error', 'pos_error'))
            from ..spe.spe_utils import _get_score_engine
            for images, targets in dataloader[phase]:
                pose, _ = spe_model.predict(images['torch'])
                tgt = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in targets.items()}
                s, p = _get_score_engine().score_host(pose['ori'], pose['pos'], tgt['ori'], tgt['pos'], want_per_image=True)
                running_avg.update(SPEUtils.metrics_from_sums(s), images['torch'].size(0))
                sums += s
                per.append(p)
            per_image = np.concatenate(per, axis=0) if per else np.zeros((0, 2), np.float32)
            _finish(rec_score, rec_error, phase, sums, per_image)
    return rec_score, rec_error
