#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference
(/root/reference, possoj/Spacecraft-Pose-Estimation-Framework) in the build container.

    python tests/golden/make_goldens.py --calibrate   # BN running stats of the synthetic init -> package data/
    python tests/golden/make_goldens.py               # all golden .npz files under tests/golden/

The reference cannot travel to the GPU box, so its outputs are frozen here.  Inputs are regenerated from seeds by
the tests (torch / numpy CPU generators are bit-reproducible on the same build), outputs are stored.
Every array below is produced by reference code (imported through oracle/ref_loader.py); nothing from the
oracle restatement or the product package computes a stored value, except the synthetic *inputs*
(spef_b200.tools.synthetic: seeds -> weights / images / targets).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import ref_loader  # noqa: E402
from spef_b200.modeling import arch  # noqa: E402
from spef_b200.tools import synthetic  # noqa: E402


class Camera:  # src/data/datasets/speed.py:18-32 (only carried around by SPEUtils)
    fx, fy, nu, nv = 0.0176, 0.0176, 1920, 1200


def build_ref_model(ref, n_ori, n_pos, pos_mode):
    data = {"x": [({"torch": torch.rand(1, 3, 240, 384)}, {})]}
    model, _ = ref.import_model(data, "mobilenet_v2_pytorch", "ursonet_pytorch", ori_mode="classification",
                                n_ori_bins=n_ori, pos_mode=pos_mode, n_pos_bins=(n_pos if pos_mode == "classification" else None))
    return model


def calibrate(ref):
    """One train()-mode pass of the REFERENCE model (BatchNorm momentum=None => running stats = batch stats)
    over torch.rand(8,3,240,384) with the raw synthetic weights; the resulting running_mean / running_var are
    committed so that synthetic_state_dict() never has to run the network."""
    model = build_ref_model(ref, 1728, 3, "regression")
    model.load_state_dict(synthetic.raw_state_dict(1728, 3))
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.reset_running_stats()
            m.momentum = None
    model.train()
    with torch.no_grad():
        model(synthetic.synthetic_images(8, seed=4242))
    model.eval()
    sd = model.state_dict()
    out = {}
    for l in arch.conv_layers():
        p = l["prefix"]
        out[p + ".mean"] = sd[p + ".1.running_mean"].numpy().astype(np.float32)
        out[p + ".var"] = sd[p + ".1.running_var"].numpy().astype(np.float32)
    os.makedirs(os.path.dirname(synthetic.BN_CALIB_PATH), exist_ok=True)
    np.savez_compressed(synthetic.BN_CALIB_PATH, **out)
    print("wrote", synthetic.BN_CALIB_PATH, os.path.getsize(synthetic.BN_CALIB_PATH), "bytes")


def load_speed_labels(n=None):
    """Real SPEED poses shipped with the reference (src/data/datasets/speed_split/valid.json)."""
    with open(os.path.join(ref_loader.REF_ROOT, "src/data/datasets/speed_split/valid.json")) as f:
        labels = json.load(f)
    q = np.array([l["q_vbs2tango"] for l in labels], np.float64)
    t = np.array([l["r_Vo2To_vbs_true"] for l in labels], np.float64)
    return (q, t) if n is None else (q[:n], t[:n])


def golden_histograms(ref):
    out = {}
    for n in (8, 12, 16):
        for delete in (False, True):
            o = ref.OrientationSoftClassification(n, 3, delete)
            tag = f"ori{n}_{'del' if delete else 'all'}"
            out[tag + "_hist"] = o.histogram
            out[tag + "_red"] = o.redundant_flags
    for n in (24, 32):
        o = ref.OrientationSoftClassification(n, 3, False)
        out[f"ori{n}_all_nbins_nred"] = np.array([o.n_bins, int(o.redundant_flags.sum())])
        out[f"ori{n}_all_hist_sample"] = o.histogram[::97]
    p = ref.PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40]))
    out["pos10_hist"] = p.histogram
    np.savez_compressed(os.path.join(OUT, "histograms.npz"), **out)


def golden_encode_decode(ref):
    q, t = load_speed_labels()
    ori = ref.OrientationSoftClassification(12, 3, False)
    ori_del = ref.OrientationSoftClassification(12, 3, True)
    pos = ref.PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40]))
    N = 32
    enc = np.stack([ori.encode(q[i]) for i in range(N)])
    enc_del = np.stack([ori_del.encode(q[i]) for i in range(N)])
    dec, hinv = ori.decode_batch(enc)
    dec_del, _ = ori_del.decode_batch(enc_del)
    penc = np.stack([pos.encode(t[i]) for i in range(N)])
    pdec = pos.decode_batch(penc)
    # round-trip statistics over all 1800 labels
    err = []
    for i in range(q.shape[0]):
        d, _ = ori.decode(ori.encode(q[i]))
        c = min(1.0, abs(float(np.dot(d.astype(np.float64), q[i]))))
        err.append(np.degrees(2 * np.arccos(c)))
    perr = [np.linalg.norm(pos.decode(pos.encode(t[i])) - t[i]) for i in range(t.shape[0])]
    np.savez_compressed(os.path.join(OUT, "encode_decode.npz"), labels_q=q, labels_t=t, enc_ori=enc, enc_ori_del=enc_del,
                        dec_ori=dec, dec_hinv=hinv, dec_ori_del=dec_del, enc_pos=penc, dec_pos=pdec,
                        roundtrip_ori_stats=np.array([np.mean(err), np.median(err), np.max(err)]),
                        roundtrip_pos_stats=np.array([np.mean(perr), np.max(perr)]))
    print("round trip ori (mean, median, max) deg:", np.mean(err), np.median(err), np.max(err), " pos:", np.mean(perr), np.max(perr))


def golden_decode_logits(ref):
    """softmax + decode of Gaussian logits at several sharpnesses and histogram sizes (config 4)."""
    cam = Camera()
    out = {}
    for n_dim in (8, 12, 16):
        su = ref.SPEUtils(cam, "classification", n_dim, 3, False, "classification", 10, 100, None)
        for sigma in (1, 3, 10):
            rs = np.random.RandomState(1000 * n_dim + sigma)
            logits = (rs.randn(8, su.orientation.n_bins) * sigma).astype(np.float32)
            plog = (rs.randn(8, 1000) * sigma).astype(np.float32)
            pose = su.last_activ({"ori_soft": logits.copy(), "pos_soft": plog.copy()})
            soft, psoft = pose["ori_soft"].copy(), pose["pos_soft"].copy()
            pose = su.decode(pose)
            tag = f"n{n_dim}_s{sigma}"
            out[tag + "_ori_soft_row0"] = soft[0]
            out[tag + "_ori"] = pose["ori"]
            out[tag + "_argmax"] = np.argmax(logits, axis=1).astype(np.int32)
            _, hinv = su.orientation.decode_batch(soft)
            out[tag + "_hinv"] = hinv
            if n_dim == 12:
                out[tag + "_pos_soft_row0"] = psoft[0]
                out[tag + "_pos"] = pose["pos"]
    # near-uniform pdf: ill-conditioned but well defined
    su = ref.SPEUtils(cam, "classification", 12, 3, False, "regression", 10, 100, None)
    a = np.sum(su.orientation.b * np.full((1728, 1, 1), 1.0 / 1728), axis=0)
    out["uniform12_eigvals"] = np.sort(np.linalg.eigvalsh(a))
    np.savez_compressed(os.path.join(OUT, "decode_logits.npz"), **out)


def golden_score(ref):
    q, t = load_speed_labels()
    q32, t32 = q.astype(np.float32), t.astype(np.float32)
    true = {"ori": q32, "pos": t32}
    out = {}
    cases = {
        "neg": {"ori": -q32, "pos": t32},
        "roll": {"ori": np.roll(q32, 1, axis=0), "pos": np.roll(t32, 1, axis=0)},
    }
    rs = np.random.RandomState(5)
    noisy_q = q32 + rs.randn(*q32.shape).astype(np.float32) * 0.05
    noisy_q /= np.linalg.norm(noisy_q, axis=1, keepdims=True)
    cases["noisy"] = {"ori": noisy_q.astype(np.float32), "pos": (t32 + rs.randn(*t32.shape).astype(np.float32) * 0.1)}
    over = q32[:16] * np.float32(1.2)  # |q.q^| = 1.2 > 1.01: the reference clamps and does NOT raise (dead code)
    for name, pred in cases.items():
        m = ref.SPEUtils.get_score({k: v.copy() for k, v in true.items()}, pred)
        out[name] = np.array([m["esa_score"], m["ori_score"], m["pos_score"], m["ori_error"], m["pos_error"]], np.float64)
        out[name + "_pred_ori"], out[name + "_pred_pos"] = pred["ori"], pred["pos"]
    m = ref.SPEUtils.get_score({"ori": q32[:16], "pos": t32[:16]}, {"ori": over, "pos": t32[:16]})
    out["over"] = np.array([m["esa_score"], m["ori_score"], m["pos_score"], m["ori_error"], m["pos_error"]], np.float64)
    np.savez_compressed(os.path.join(OUT, "score.npz"), **out)
    print("score roll:", out["roll"], " neg:", out["neg"])


def golden_network(ref):
    """Reference ModelWrapper forward (FP32, CPU) on the calibrated synthetic weights and seeded images."""
    out = {}
    x = synthetic.synthetic_images(4)
    for tag, n_pos, mode in (("murso", 3, "regression"), ("mursop", 1000, "classification")):
        model = build_ref_model(ref, 1728, n_pos, mode)
        model.load_state_dict(synthetic.synthetic_state_dict(1728, n_pos))
        model.eval()
        with torch.no_grad():
            ori, pos = model(x)
            feats = model.features.features[0](x)
        out[tag + "_ori"], out[tag + "_pos"] = ori.numpy(), pos.numpy()
        if tag == "murso":
            out["stem_out_b0_c0to3"] = feats[0, :4].numpy()  # a slice of the stem activation (layout check)
            # activation statistics per feature index (documents that the init is non-degenerate)
            stats, h = [], x
            with torch.no_grad():
                for m in model.features.features:
                    h = m(h)
                    stats.append([float(h.mean()), float(h.std())])
            out["feature_stats"] = np.array(stats)
    np.savez_compressed(os.path.join(OUT, "network.npz"), **out)
    print("logits std", out["murso_ori"].std(), "absmax", np.abs(out["murso_ori"]).max(), "pos", out["murso_pos"][0])


def golden_evaluation(ref):
    """Reference evaluation() with SPETorch on CPU over a 3-batch synthetic loader (FP32)."""
    cam = Camera()
    su = ref.SPEUtils(cam, "classification", 12, 3, False, "regression", 10, 100, None)
    model = build_ref_model(ref, 1728, 3, "regression")
    model.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
    spe = ref.SPETorch(model, torch.device("cpu"), su)
    loader = synthetic.SyntheticLoader(10, 4)
    rec_score, rec_error = ref.evaluation(spe, {"valid": loader}, su, ("valid",))
    poses = [spe.predict(b[0]["torch"])[0] for b in loader]
    np.savez_compressed(os.path.join(OUT, "evaluation.npz"),
                        score=np.array([rec_score["valid"][k][0] for k in ("ori", "pos", "esa")], np.float64),
                        error=np.array([rec_error["valid"][k][0] for k in ("ori", "pos", "ori_std", "pos_std", "ori_mad", "pos_mad")], np.float64),
                        pred_ori=np.concatenate([p["ori"] for p in poses]), pred_pos=np.concatenate([p["pos"] for p in poses]))
    print("evaluation:", rec_score, rec_error)


def golden_temporal(ref):
    """Reference TemporalPDF / decode / sign-continuity trace (Inference.predict 'Adaptative', inference.py:131-180)
    driven with synthetic logits: encoded pdfs along a D-SPEED-like constant-rate trajectory
    (create_dspeed.py:299-311) + noise, an outlier frame and a sign-flip, so every branch of the filter fires."""
    cam = Camera()
    su = ref.SPEUtils(cam, "classification", 12, 3, False, "classification", 10, 100, None)
    T = 16
    rs = np.random.RandomState(11)
    q0 = np.array([0.0, -0.7071, 0.7071, 0.0])
    q0 /= np.linalg.norm(q0)
    t0 = np.array([-7.0, -4.5, 30.0])
    ori_logits, pos_logits = [], []
    for k in range(T):
        ang = np.deg2rad(0.9 * k)
        dq = np.array([np.cos(ang / 2), np.sin(ang / 2) * 0.6, np.sin(ang / 2) * 0.64, np.sin(ang / 2) * 0.48])
        q = np.array([dq[0] * q0[0] - dq[1:] @ q0[1:], *(dq[0] * q0[1:] + q0[0] * dq[1:] + np.cross(dq[1:], q0[1:]))])
        q /= np.linalg.norm(q)
        t = t0 + k * np.array([0.0048, 0.0032, -0.016]) * 10
        if k == 9:  # outlier frame
            q = np.array([0.5, 0.5, -0.5, 0.5])
            t = np.array([5.0, 5.0, 10.0])
        po = su.orientation.encode(q).astype(np.float64)
        pp = su.position.encode(t).astype(np.float64)
        lo = np.log(po + 1e-6) + rs.randn(po.size) * 0.3
        lp = np.log(pp + 1e-6) + rs.randn(pp.size) * 0.3
        ori_logits.append(lo.astype(np.float32))
        pos_logits.append(lp.astype(np.float32))
    ori_logits, pos_logits = np.stack(ori_logits), np.stack(pos_logits)

    # drive the reference's own Inference object with a stub engine that returns these logits through the
    # reference's post-processing (SPETorch.predict = last_activ + decode, spe_torch.py:75-76)
    class StubEngine:
        def __init__(self):
            self.k = 0

        def predict(self, image):
            pose = {"ori_soft": ori_logits[self.k:self.k + 1].copy(), "pos_soft": pos_logits[self.k:self.k + 1].copy()}
            self.k += 1
            pose = su.last_activ(pose)
            pose = su.decode(pose)
            return pose, 0.0

    inf = ref.Inference.__new__(ref.Inference)
    inf.model, inf.inference_device, inf.spe_utils = None, "cpu_host", su
    inf.inference_engine = StubEngine()
    inf.prev_still_ori = inf.prev_video_ori = None
    inf.pdf_adapt_ori = ref.TemporalPDF(n=0.8, alpha=16.49, distance_metric="l2")  # inference.py:38-39
    inf.pdf_adapt_pos = ref.TemporalPDF(n=0.5, alpha=48.64, distance_metric="l2")
    inf.ssh_jetson, inf.img_size = None, None
    rec = {k: [] for k in ("still_ori", "still_pos", "video_ori", "video_pos", "ori_distance", "pos_distance",
                           "video_ori_soft_row", "video_pos_soft_row")}
    img = torch.zeros(1, 3, 8, 8)
    for k in range(T):
        still, _, video = inf.predict(img, "Adaptative")
        rec["still_ori"].append(still["ori"]); rec["still_pos"].append(still["pos"])
        rec["video_ori"].append(video["ori"]); rec["video_pos"].append(video["pos"])
        rec["ori_distance"].append(video["ori_distance"]); rec["pos_distance"].append(video["pos_distance"])
        rec["video_ori_soft_row"].append(video["ori_soft"][::16]); rec["video_pos_soft_row"].append(video["pos_soft"][::16])
    np.savez_compressed(os.path.join(OUT, "temporal.npz"), ori_logits=ori_logits, pos_logits=pos_logits,
                        **{k: np.array(v) for k, v in rec.items()})
    print("temporal distances:", np.array(rec["ori_distance"]).round(4))


def golden_resize(ref):
    """Input side: the reference's own transform, transforms.Compose([Resize(img_size), ToTensor()])
    (src/data/datasets/speed.py:59-62) applied to Image.fromarray(frame).convert("RGB") as SPEDataset.__getitem__ does
    (src/data/utils.py:215-226).  Frames are regenerated from seeds by the tests; the stored values are the 8-bit pixels
    behind ToTensor's output (checked here to be exactly out * 255) plus two float32 samples of the tensor itself."""
    from PIL import Image
    from torchvision import transforms
    cases = [  # name, frames kwargs, img_size
        ("speed_1200x1920", dict(batch=2, height=1200, width=1920, channels=1, seed=11, kind="speed"), (240, 384)),
        ("noise_1200x1920", dict(batch=1, height=1200, width=1920, channels=1, seed=12, kind="noise"), (240, 384)),
        ("noise_1200x1920_sq", dict(batch=1, height=1200, width=1920, channels=1, seed=13, kind="noise"), (240, 240)),
        ("rgb_480x640", dict(batch=1, height=480, width=640, channels=3, seed=14, kind="noise"), (240, 384)),
        ("up_123x257", dict(batch=1, height=123, width=257, channels=1, seed=15, kind="noise"), (240, 384)),
        ("odd_601x997_rgb", dict(batch=1, height=601, width=997, channels=3, seed=16, kind="speed"), (240, 384)),
        ("same_240x384", dict(batch=1, height=240, width=384, channels=1, seed=17, kind="noise"), (240, 384)),
    ]
    out = {"cases": json.dumps([[n, k, list(s)] for n, k, s in cases])}
    for name, kw, size in cases:
        frames = synthetic.synthetic_frames(**kw)
        tf = transforms.Compose([transforms.Resize(size), transforms.ToTensor()])
        res = []
        for fr in frames:
            t = tf(Image.fromarray(fr).convert("RGB"))
            u8 = torch.round(t * 255).to(torch.uint8)
            assert torch.equal(u8.float() / 255, t)
            res.append(u8.numpy())
        res = np.stack(res)
        if kw["channels"] == 1:
            assert (res[:, 0] == res[:, 1]).all() and (res[:, 0] == res[:, 2]).all()
            res = res[:, :1]
        out[name] = res
        print(name, res.shape, res.mean())
    t = transforms.Compose([transforms.Resize((240, 384)), transforms.ToTensor()])(
        Image.fromarray(synthetic.synthetic_frames(**cases[0][1])[0]).convert("RGB"))
    out["speed_1200x1920_f32_row100"] = t[:, 100, :].numpy()
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calibrate", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    ref = ref_loader.load()
    torch.set_num_threads(8)
    if args.calibrate:
        calibrate(ref)
        return
    steps = {"histograms": golden_histograms, "encode_decode": golden_encode_decode, "decode_logits": golden_decode_logits,
             "score": golden_score, "network": golden_network, "evaluation": golden_evaluation, "temporal": golden_temporal, "resize": golden_resize}
    for name, fn in steps.items():
        if args.only and name != args.only:
            continue
        print("==", name)
        fn(ref)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
