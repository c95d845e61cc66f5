"""N > 1 host logic on CPU: two gloo ranks each hold a shard of the evaluation and combine the eight ESA
accumulators with the single SUM all-reduce of the path; the result must equal the one-rank evaluation
(src/tools/utils.py:84-95 is linear, so sum_b(mean_b n_b)/sum_b n_b == global per-image mean)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_rank_results(rank, world, n_images=37, batch=4):
    """What rank `rank` would hold after its shard: float64 sums + per-image errors (deterministic numbers that stand
    in for kernel outputs; the reduction logic is what is under test)."""
    from spef_b200.tools import synthetic
    loader = synthetic.SyntheticLoader(n_images, batch, img_size=(8, 8), rank=rank, world=world)
    sums, per = np.zeros(8), []
    for images, tgt in loader:
        q = tgt["ori"].numpy().astype(np.float64)
        t = tgt["pos"].numpy().astype(np.float64)
        e_q, e_t = np.abs(q[:, 0]) * 2.0, np.abs(t[:, 0])
        sums[0] += e_q.sum(); sums[1] += (e_t / np.linalg.norm(t, axis=1)).sum(); sums[2] += e_t.sum(); sums[3] += len(q)
        per.append(np.stack([np.degrees(e_q), e_t], axis=1).astype(np.float32))
    return sums, np.concatenate(per, 0)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spef_b200.tools.evaluation import reduce_eval_sums, gather_per_image, _finish
    sums, per = _fake_rank_results(rank, world)
    rec_score = {"valid": {'ori': [], 'pos': [], 'esa': []}}
    rec_error = {"valid": {'ori': [], 'pos': [], 'ori_std': [], 'pos_std': [], 'ori_mad': [], 'pos_mad': []}}
    _finish(rec_score, rec_error, "valid", sums, per)
    red = reduce_eval_sums(sums)
    allper = gather_per_image(per)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), red=red, n=allper.shape[0],
             score=[rec_score["valid"][k][0] for k in ("ori", "pos", "esa")],
             error=[rec_error["valid"][k][0] for k in ("ori", "pos", "ori_std", "pos_std", "ori_mad", "pos_mad")])
    dist.destroy_process_group()


def test_two_rank_evaluation_equals_one_rank(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from spef_b200.tools.evaluation import _finish
    sums1, per1 = _fake_rank_results(0, 1)
    rec_score = {"valid": {'ori': [], 'pos': [], 'esa': []}}
    rec_error = {"valid": {'ori': [], 'pos': [], 'ori_std': [], 'pos_std': [], 'ori_mad': [], 'pos_mad': []}}
    _finish(rec_score, rec_error, "valid", sums1, per1)
    want_score = [rec_score["valid"][k][0] for k in ("ori", "pos", "esa")]
    want_error = [rec_error["valid"][k][0] for k in ("ori", "pos", "ori_std", "pos_std", "ori_mad", "pos_mad")]
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        np.testing.assert_allclose(g["red"], sums1, rtol=1e-12)  # float64 sums: only the summation order differs
        assert int(g["n"]) == 37 == int(g["red"][3])
        np.testing.assert_allclose(g["score"], want_score, rtol=1e-6)
        np.testing.assert_allclose(g["error"], want_error, rtol=1e-6)


def test_loader_sharding_covers_every_image_once():
    from spef_b200.tools import synthetic
    full = synthetic.SyntheticLoader(21, 4, img_size=(8, 8))
    shards = [synthetic.SyntheticLoader(21, 4, img_size=(8, 8), rank=r, world=4) for r in range(4)]
    assert sum(len(s) for s in shards) == len(full) == 6
    got = sorted(float(t["pos"][0, 2]) for s in shards for _, t in s)
    assert got == sorted(float(t["pos"][0, 2]) for _, t in full)
