"""GPU parity: softmax / decode / score / temporal kernels (through the C ABI) against the CPU oracle and the
reference goldens.  Tolerances are the north-star ones: decoded quaternions <= 0.05 deg, positions <= 1e-4 relative,
argmax / bin indices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import spef_oracle as O

pytestmark = pytest.mark.gpu

QUAT_TOL_DEG = 0.05
POS_RTOL = 1e-4


@pytest.fixture(scope="module", params=["stream", "per_image"])
def post(request):
    """spef_decode_ori runs decode_ori_stream_kernel by default; SPEF_DECODE_STREAM=0 keeps the first-generation
    warp-per-image kernel (also the path of rows that are not 16-byte aligned) -- every test below runs on both."""
    import os
    from spef_b200.engine import Engine
    if request.param == "per_image":
        os.environ["SPEF_DECODE_STREAM"] = "0"
    try:
        return Engine(32, 32, 8, 3, False, "fp32", 1)
    finally:
        os.environ.pop("SPEF_DECODE_STREAM", None)


def _engine_for(post, n_dim, delete=False):
    hist, _ = O.ori_histogram(n_dim, delete)
    post.set_ori_histogram(hist)
    post.set_pos_histogram(O.pos_histogram(10))
    return hist


@pytest.mark.parametrize("n_dim", [8, 12, 16])
@pytest.mark.parametrize("sigma", [1, 3, 10])
def test_softmax_decode_vs_reference_golden(post, golden, n_dim, sigma):
    g = golden("decode_logits")
    hist = _engine_for(post, n_dim)
    rs = np.random.RandomState(1000 * n_dim + sigma)
    logits = (rs.randn(8, hist.shape[0]) * sigma).astype(np.float32)
    plog = (rs.randn(8, 1000) * sigma).astype(np.float32)
    tag = f"n{n_dim}_s{sigma}"
    out = post.decode_ori_host(logits, is_logits=True, want_soft=True, want_hinv=True, want_argmax=True)
    assert not out["flags"].any()
    assert O.quat_angle_deg(out["quat"], g[tag + "_ori"]).max() <= QUAT_TOL_DEG
    np.testing.assert_array_equal(out["argmax"], g[tag + "_argmax"])  # bit-exact
    np.testing.assert_allclose(out["soft"][0], g[tag + "_ori_soft_row0"], rtol=3e-6, atol=1e-30)
    np.testing.assert_allclose(np.linalg.norm(out["quat"], axis=1), 1.0, atol=1e-6)
    # inv(A): conditioned like A; compare relative to the largest entry
    assert np.abs(out["hinv"] - g[tag + "_hinv"]).max() <= 1e-3 * np.abs(g[tag + "_hinv"]).max()
    # device-pointer entry point gives the same answer as the host-buffer one
    dev = post.decode_ori(torch.from_numpy(logits), is_logits=True, want_soft=True, want_hinv=True, want_argmax=True)
    np.testing.assert_array_equal(dev["quat"].cpu().numpy(), out["quat"])
    np.testing.assert_array_equal(dev["soft"].cpu().numpy(), out["soft"])
    if n_dim == 12:
        pout = post.decode_pos_host(plog, is_logits=True, want_soft=True)
        np.testing.assert_allclose(pout["pos"], g[tag + "_pos"], rtol=POS_RTOL)
        np.testing.assert_allclose(pout["soft"][0], g[tag + "_pos_soft_row0"], rtol=3e-6, atol=1e-30)


def test_decode_of_encoded_labels_vs_reference_golden(post, golden):
    g = golden("encode_decode")
    for delete, key in ((False, ""), (True, "_del")):
        _engine_for(post, 12, delete)
        out = post.decode_ori_host(g["enc_ori" + key], is_logits=False, want_hinv=(not delete))
        assert O.quat_angle_deg(out["quat"], g["dec_ori" + key]).max() <= QUAT_TOL_DEG
        # round trip back to the real SPEED label: within the histogram's resolution (max 5.17 deg)
        assert O.quat_angle_deg(out["quat"], g["labels_q"][:out["quat"].shape[0]]).max() < 5.2
    pout = post.decode_pos_host(g["enc_pos"], is_logits=False)
    np.testing.assert_allclose(pout["pos"], g["dec_pos"], rtol=POS_RTOL, atol=1e-5)


@pytest.mark.parametrize("n_dim", [8, 12, 16, 24, 32])
def test_decode_sweep_vs_oracle(post, n_dim):
    """BASELINE config 4: bins per axis 8..32 (512 .. 32768 bins), Gaussian logits and peaked pdfs."""
    hist = _engine_for(post, n_dim)
    n = hist.shape[0]
    rs = np.random.RandomState(n_dim)
    logits = (rs.randn(6, n) * 3).astype(np.float32)
    soft = O.softmax(logits)
    want, _ = O.ori_decode_batch(soft, hist)
    out = post.decode_ori_host(logits, is_logits=True, want_soft=True, want_argmax=True)
    assert O.quat_angle_deg(out["quat"], want).max() <= QUAT_TOL_DEG
    np.testing.assert_array_equal(out["argmax"], np.argmax(logits, axis=1))
    np.testing.assert_allclose(out["soft"], soft, rtol=3e-6, atol=1e-30)
    out2 = post.decode_ori_host(soft, is_logits=False)  # pdf input path
    assert O.quat_angle_deg(out2["quat"], want).max() <= QUAT_TOL_DEG


def test_decode_edge_cases(post):
    hist = _engine_for(post, 12)
    n = hist.shape[0]
    # ties: np.argmax returns the first maximum
    z = np.zeros((3, n), np.float32)
    z[1, [7, 900]] = 2.0
    z[2, -1] = 1.0
    out = post.decode_ori_host(z, is_logits=True, want_argmax=True)
    np.testing.assert_array_equal(out["argmax"], [0, 7, n - 1])
    # single image, one-hot pdf -> that bin's quaternion (up to sign)
    p = np.zeros((1, n), np.float32)
    p[0, 123] = 1.0
    out = post.decode_ori_host(p, is_logits=False)
    assert O.quat_angle_deg(out["quat"][0], hist[123]) < 1e-3
    # NaN guard (classification_utils.py:134-135) and zero-sum guard (:253-254) surface as flags / ValueError
    bad = np.zeros((2, n), np.float32)
    bad[1, 5] = np.nan
    out = post.decode_ori_host(bad, is_logits=True)
    assert out["flags"][0] == 0 and out["flags"][1] & 1
    from spef_b200.spe import OrientationSoftClassification, PositionSoftClassification
    with pytest.raises(ValueError, match="Error during orientation decoding"):
        OrientationSoftClassification(12, 3, False).decode_batch(bad)
    with pytest.raises(ValueError, match="sum is zero"):
        PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40])).decode(np.zeros(1000, np.float32))


def test_score_vs_reference_golden(post, golden):
    g, ed = golden("score"), golden("encode_decode")
    q32, t32 = ed["labels_q"].astype(np.float32), ed["labels_t"].astype(np.float32)
    from spef_b200.spe import SPEUtils
    for name in ("neg", "roll", "noisy"):
        pred = {"ori": g[name + "_pred_ori"], "pos": g[name + "_pred_pos"]}
        m = SPEUtils.get_score({"ori": q32, "pos": t32}, pred)
        got = np.array([m[k] for k in ("esa_score", "ori_score", "pos_score", "ori_error", "pos_error")], np.float64)
        # reference: float32 pairwise means; ours: float64 sums of the same float32 per-image terms
        np.testing.assert_allclose(got, g[name], rtol=1e-5, atol=1e-7)
        sums, per = post.score_host(pred["ori"], pred["pos"], q32, t32, want_per_image=True)
        eo, ep = O.per_image_errors({"ori": q32, "pos": t32}, pred)
        np.testing.assert_allclose(per[:, 1], ep, rtol=1e-6)          # sqrt/add/mul are exactly rounded on both sides
        np.testing.assert_allclose(per[:, 0], eo, rtol=1e-5, atol=2e-3)  # acosf vs NumPy arccos: <= 2 ulp of the cosine
        assert sums[3] == q32.shape[0] and sums[4] == 0 and sums[5] == 0
        want64 = np.array([np.sum(np.radians(eo.astype(np.float64))), 0, np.sum(ep.astype(np.float64))])
        np.testing.assert_allclose(sums[[0, 2]], want64[[0, 2]], rtol=1e-5)
    # |q.q^| = 1.2: clamped, counted as a diagnostic, never raised (the reference's check is dead code)
    m = SPEUtils.get_score({"ori": q32[:16], "pos": t32[:16]}, {"ori": q32[:16] * np.float32(1.2), "pos": t32[:16]})
    assert m["ori_score"] == 0.0 == g["over"][1]
    sums, _ = post.score_host(q32[:16] * np.float32(1.2), t32[:16], q32[:16], t32[:16])
    assert sums[4] == 16


def test_speutils_facade_vs_oracle():
    from spef_b200.spe import SPEUtils
    su = SPEUtils(None, 'classification', 12, 3, False, 'classification', 10, 100, None)
    rs = np.random.RandomState(3)
    logits = (rs.randn(5, 1728) * 4).astype(np.float32)
    plog = (rs.randn(5, 1000) * 4).astype(np.float32)
    pose = su.last_activ({"ori_soft": logits.copy(), "pos_soft": plog.copy()})
    np.testing.assert_allclose(pose["ori_soft"], O.softmax(logits), rtol=3e-6, atol=1e-30)
    np.testing.assert_allclose(pose["pos_soft"], O.softmax(plog), rtol=3e-6, atol=1e-30)
    pose = su.decode(pose)
    want_q, want_h = O.ori_decode_batch(O.softmax(logits), O.ori_histogram(12)[0])
    assert pose["ori"].dtype == np.float32 and pose["ori"].shape == (5, 4)
    assert O.quat_angle_deg(pose["ori"], want_q).max() <= QUAT_TOL_DEG
    np.testing.assert_allclose(pose["pos"], O.pos_decode_batch(O.softmax(plog), O.pos_histogram(10)), rtol=POS_RTOL)
    q, h = su.orientation.decode(pose["ori_soft"][0])  # (q, h_inv) tuple like the reference
    assert q.shape == (4,) and h.shape == (4, 4)
    assert np.abs(h - want_h[0]).max() <= 1e-3 * np.abs(want_h[0]).max()


def test_temporal_trace_vs_reference_golden(golden):
    """Inference.predict(..., 'Adaptative') after the network, on the golden logits stream (16 frames with an outlier)."""
    from spef_b200.engine import Engine
    g = golden("temporal")
    eng = Engine(32, 32, 1728, 1000, True, "fp32", 2)
    eng.set_ori_histogram(O.ori_histogram(12)[0])
    eng.set_pos_histogram(O.pos_histogram(10))
    eng.temporal_reset(2)  # two identical streams: stream independence
    for k in range(g["ori_logits"].shape[0]):
        lo = np.stack([g["ori_logits"][k]] * 2)
        lp = np.stack([g["pos_logits"][k]] * 2)
        out = {n: v.cpu().numpy() for n, v in eng.temporal_step_logits(lo, lp).items()}
        for s in range(2):
            assert O.quat_angle_deg(out["still_quat"][s], g["still_ori"][k]) <= QUAT_TOL_DEG
            assert O.quat_angle_deg(out["video_quat"][s], g["video_ori"][k]) <= QUAT_TOL_DEG
            if k > 0:  # after the first frame the sign is fixed by continuity with the previous frame
                assert np.sign(np.dot(out["still_quat"][s], prev_still)) == np.sign(np.dot(g["still_ori"][k], g["still_ori"][k - 1]))
                assert np.dot(out["video_quat"][s], prev_video) * np.dot(g["video_ori"][k], g["video_ori"][k - 1]) > 0
            np.testing.assert_allclose(out["still_pos"][s], g["still_pos"][k], rtol=POS_RTOL)
            np.testing.assert_allclose(out["video_pos"][s], g["video_pos"][k], rtol=POS_RTOL)
            np.testing.assert_allclose(out["ori_distance"][s], g["ori_distance"][k], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(out["pos_distance"][s], g["pos_distance"][k], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(out["video_ori_soft"][s][::16], g["video_ori_soft_row"][k], rtol=1e-4, atol=1e-10)
            np.testing.assert_allclose(out["video_pos_soft"][s][::16], g["video_pos_soft_row"][k], rtol=1e-4, atol=1e-10)
        prev_still, prev_video = out["still_quat"][0], out["video_quat"][0]
    # reset restarts the filter: next frame is a "first frame" again (distance 0)
    eng.temporal_reset(2)
    out = eng.temporal_step_logits(np.stack([g["ori_logits"][3]] * 2), np.stack([g["pos_logits"][3]] * 2))
    assert float(out["ori_distance"].abs().max()) == 0.0


def test_temporal_pdf_facade_vs_oracle():
    from spef_b200.temporal import TemporalPDF
    rs = np.random.RandomState(9)
    ours, ref = TemporalPDF(n=0.8, alpha=16.49), O.TemporalPDF(0.8, 16.49)
    base = rs.rand(1728).astype(np.float32)
    for k in range(6):
        p = (base + 0.2 * rs.rand(1728) * (k % 3)).astype(np.float32)
        a, da = ours.update_pdf(p)
        b, db = ref.update_pdf(p)
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(da, db, rtol=1e-4, atol=1e-8)
    ours.reset()
    _, d = ours.update_pdf(base)
    assert d == 0.0 and ours.previous_pdf is not None


def test_decode_large_batch_properties(post):
    """Size-independent properties at a batch the oracle could not finish in seconds (B = 8192, 1728 bins):
    shift invariance of softmax, permutation equivariance, unit norm, determinism."""
    hist = _engine_for(post, 12)
    g = torch.Generator().manual_seed(5)
    logits = torch.randn((8192, 1728), generator=g) * 5
    a = post.decode_ori(logits, is_logits=True, want_soft=True, want_argmax=True)
    b = post.decode_ori(logits + 3.25, is_logits=True, want_argmax=True)  # exact in f32 for |z| < 2^18: same max-subtracted values +- 1 ulp
    perm = torch.randperm(8192, generator=g)
    c = post.decode_ori(logits[perm], is_logits=True, want_argmax=True)
    qa, qb, qc = a["quat"].cpu().numpy(), b["quat"].cpu().numpy(), c["quat"].cpu().numpy()
    assert np.abs(np.linalg.norm(qa, axis=1) - 1).max() < 1e-6
    assert O.quat_angle_deg(qa, qb).max() < 0.01
    np.testing.assert_array_equal(qc, qa[perm.numpy()])
    np.testing.assert_array_equal(c["argmax"].cpu().numpy(), a["argmax"].cpu().numpy()[perm.numpy()])
    np.testing.assert_array_equal(a["argmax"].cpu().numpy(), logits.argmax(1).numpy())
    np.testing.assert_allclose(a["soft"].sum(1).cpu().numpy(), 1.0, atol=1e-5)
    # spot-check 8 images against the oracle
    idx = [0, 1, 1000, 4095, 4096, 6000, 8190, 8191]
    want, _ = O.ori_decode_batch(O.softmax(logits[idx].numpy()), hist)
    assert O.quat_angle_deg(qa[idx], want).max() <= QUAT_TOL_DEG


def test_encode_batch_vs_reference_golden(golden):
    """spef_encode_ori / spef_encode_pos (label side) against the pdfs the unmodified reference encoded for 32 real
    SPEED labels (tests/golden/encode_decode.npz: all bins / unused bins deleted / position); float64 in the kernel like NumPy."""
    from spef_b200.spe.classification_utils import OrientationSoftClassification, PositionSoftClassification
    g = golden("encode_decode")
    q, t = g["labels_q"][:32], g["labels_t"][:32]
    osc = OrientationSoftClassification(12, 3, False)
    got = osc.encode_batch(q)
    assert got.shape == g["enc_ori"].shape and got.dtype == np.float32
    np.testing.assert_allclose(got, g["enc_ori"], rtol=2e-6, atol=1e-30)
    np.testing.assert_allclose(got.sum(1), 1.0, atol=1e-6)
    assert (got[:, osc.redundant_flags] == 0).all()
    np.testing.assert_allclose(OrientationSoftClassification(12, 3, True).encode_batch(q), g["enc_ori_del"], rtol=2e-6, atol=1e-30)
    psc = PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40]))
    np.testing.assert_allclose(psc.encode_batch(t), g["enc_pos"], rtol=2e-6, atol=1e-30)
    # other histogram sizes against the oracle
    hist, red = O.ori_histogram(16, False)
    want = np.stack([O.ori_encode(qq, hist, red, 16, 3, False) for qq in q[:8]])
    np.testing.assert_allclose(OrientationSoftClassification(16, 3, False).encode_batch(q[:8]), want, rtol=2e-6, atol=1e-30)
    # encode -> decode round trip equals the reference's decoded labels
    q_dec, _ = osc.decode_batch(got)
    assert O.quat_angle_deg(q_dec, g["dec_ori"]).max() <= 0.05
    with pytest.raises(ValueError):
        osc.encode_batch(np.full((1, 4), np.nan))


@pytest.mark.parametrize("n", [1, 2, 7, 1800, 4097])
def test_error_stats_vs_numpy(n):
    """spef_error_stats = np.mean / np.std / np.median / mad() of evaluation.py:16-32 on a float32 column."""
    from spef_b200.engine import Engine
    eng = Engine(32, 32, 8, 3, False, "fp32", 1)
    rng = np.random.default_rng(n)
    x = np.abs(rng.standard_normal((n, 2))).astype(np.float32) * np.array([30.0, 0.5], np.float32)
    x[::5] = x[0]                                    # duplicates
    if n > 10:
        x[3, 0] = 1e4                                # an outlier
    for col in (0, 1):
        s = eng.error_stats(torch.from_numpy(x).cuda(), col)
        c = x[:, col]
        assert s["median"] == float(np.median(c))                       # exact: radix select + float32 mean of the middle pair
        assert s["mad"] == float(O.mad(list(c)))
        np.testing.assert_allclose(s["mean"], c.astype(np.float64).mean(), rtol=1e-12)
        np.testing.assert_allclose(s["std"], c.astype(np.float64).std(), rtol=1e-10)
        np.testing.assert_allclose(s["std"], np.std(c), rtol=2e-5)     # NumPy's float32 reduction


def test_evaluation_device_stats_matches_host():
    from spef_b200.modeling import import_model
    from spef_b200.spe import SPEB200, SPEUtils
    from spef_b200.tools import evaluation, synthetic
    su = SPEUtils(None, 'classification', 12, 3, False, 'regression', 10, 100, None)
    loader = synthetic.SyntheticLoader(10, 4)
    model, _ = import_model({"valid": loader}, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification',
                            n_ori_bins=su.orientation.n_bins, pos_mode='regression', precision="bf16")
    model.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
    spe = SPEB200(model, torch.device("cuda:0"), su)
    s_host, e_host = evaluation(spe, {"valid": loader}, su, ("valid",))
    s_dev, e_dev = evaluation(spe, {"valid": loader}, su, ("valid",), device_stats=True)
    assert s_host == s_dev
    for k in ("ori", "pos"):
        assert e_host["valid"][k] == e_dev["valid"][k]
    for k in ("ori_std", "pos_std"):
        np.testing.assert_allclose(e_dev["valid"][k], e_host["valid"][k], rtol=2e-5)
    for k in ("ori_mad", "pos_mad"):
        np.testing.assert_allclose(e_dev["valid"][k], e_host["valid"][k], rtol=1e-6)


# ---- the large-batch (streaming) decode kernel: csrc/decode_stream.cuh ---------------------------------
@pytest.fixture(scope="module")
def stream_post():
    """An engine on the streaming decode kernel (the default)."""
    from spef_b200.engine import Engine
    return Engine(32, 32, 8, 3, False, "fp32", 1)


@pytest.mark.parametrize("n_dim", [8, 12, 16, 24])
def test_stream_decode_vs_oracle(stream_post, n_dim):
    """Same gates as the per-image kernel: <= 0.05 deg on identical logits / pdfs, argmax bit-exact, inv(A) 1e-3.
    n_dim 16 and 24 take the chunked-table path (n > 2048 bins), 12 and 24 have a ragged last 1024-bin step."""
    hist, _ = O.ori_histogram(n_dim)
    stream_post.set_ori_histogram(hist)
    n = hist.shape[0]
    rs = np.random.RandomState(77 + n_dim)
    B = 70   # small-batch configuration (8 warps x 4 ring slots): 9 groups, the last one ragged
    for sigma in (1.0, 3.0, 10.0):
        logits = (rs.randn(B, n) * sigma).astype(np.float32)
        out = stream_post.decode_ori(torch.from_numpy(logits), is_logits=True, want_soft=True, want_hinv=True, want_argmax=True)
        soft = O.softmax(logits)
        want_q, want_h = O.ori_decode_batch(soft, hist)
        assert not out["flags"].cpu().numpy().any()
        assert O.quat_angle_deg(out["quat"].cpu().numpy(), want_q).max() <= QUAT_TOL_DEG
        np.testing.assert_array_equal(out["argmax"].cpu().numpy(), logits.argmax(1))
        np.testing.assert_allclose(out["soft"].cpu().numpy(), soft, rtol=5e-6, atol=1e-30)
        # inv(A) amplifies the float32 rounding of the softmax weights (shared with the reference's own float32 pdf) by cond(A):
        # gate 1e-3 for well conditioned images, cond * 2e-6 beyond (one-hot pdfs at sigma = 10 reach cond 1e9)
        hv = out["hinv"].cpu().numpy()
        A = np.einsum("ib,bj,bk->ijk", soft.astype(np.float64), hist, hist)
        cond = np.linalg.cond(A)
        rel = np.abs(hv - want_h).reshape(B, -1).max(1) / np.abs(want_h).reshape(B, -1).max(1)
        assert (rel <= np.maximum(1e-3, 2e-6 * cond)).all(), (rel.max(), cond[rel.argmax()])
        assert (cond < 1e4).sum() >= (0 if sigma >= 10 else 10)
        # pdf input (the temporal path decodes filtered pdfs)
        out2 = stream_post.decode_ori(torch.from_numpy(soft), is_logits=False, want_argmax=True)
        assert O.quat_angle_deg(out2["quat"].cpu().numpy(), want_q).max() <= QUAT_TOL_DEG
        np.testing.assert_array_equal(out2["argmax"].cpu().numpy(), soft.argmax(1))


def test_stream_decode_edge_cases(stream_post, golden):
    hist, _ = O.ori_histogram(12)
    stream_post.set_ori_histogram(hist)
    n = hist.shape[0]
    # ties: first maximum wins (np.argmax); a maximum in the last bin; constant rows
    z = np.zeros((9, n), np.float32)
    z[0, [5, 900, 1700]] = 4.0
    z[1, n - 1] = 2.0
    z[2, 1024] = 1.0            # first bin of the second 1024-bin step
    z[3, [1023, 1024]] = 3.0
    z[4] = -50.0
    z[5, 17] = 80.0             # one-hot after softmax
    z[6] = np.linspace(-20, 20, n)
    z[7, 1] = np.nan            # -> flagged like the reference's ValueError
    out = stream_post.decode_ori(torch.from_numpy(z), is_logits=True, want_argmax=True)
    am = out["argmax"].cpu().numpy()
    np.testing.assert_array_equal(am[:7], z[:7].argmax(1))
    fl = out["flags"].cpu().numpy()
    assert fl[7] == 1 and not fl[:7].any() and not fl[8]
    q = out["quat"].cpu().numpy()
    assert np.isnan(q[7]).all()
    want_q, _ = O.ori_decode_batch(O.softmax(z[[0, 1, 2, 3, 6]]), hist)
    assert O.quat_angle_deg(q[[0, 1, 2, 3, 6]], want_q).max() <= QUAT_TOL_DEG
    assert O.quat_angle_deg(q[5], hist[17].astype(np.float32)) <= QUAT_TOL_DEG   # one-hot pdf: A = q q^T (singular: the oracle's inv() raises)
    # encoded SPEED labels from the reference
    g = golden("encode_decode")
    out = stream_post.decode_ori(torch.from_numpy(g["enc_ori"]), is_logits=False, want_hinv=True)
    assert O.quat_angle_deg(out["quat"].cpu().numpy(), g["dec_ori"]).max() <= QUAT_TOL_DEG


def test_stream_decode_full_size_matches_per_image_kernel(post, stream_post):
    """BASELINE configs[3] size (1728 bins, 77 672 images: every CTA loops, 32-image solve batches plus a remainder):
    the default engine dispatches to the streaming kernel by itself; results against the per-image kernel on the same
    device buffer, plus slot independence and determinism."""
    hist, _ = O.ori_histogram(12)
    post.set_ori_histogram(hist)
    stream_post.set_ori_histogram(hist)
    g = torch.Generator().manual_seed(5)
    B = 77672
    logits = (torch.randn((B, 1728), generator=g) * 3).cuda()
    a = stream_post.decode_ori(logits, is_logits=True, want_argmax=True)            # 16 warps x 2 ring slots + L2 prefetch
    b = stream_post.decode_ori(logits, is_logits=True, want_argmax=True)
    assert torch.equal(a["quat"], b["quat"])
    assert torch.equal(a["argmax"].long(), logits.argmax(1))
    assert not a["flags"].any()
    qa = a["quat"].cpu().numpy()
    ref = np.concatenate([post.decode_ori(logits[i:i + 4096], is_logits=True)["quat"].cpu().numpy() for i in range(0, B, 4096)])  # 4096-image slices, both kernels
    assert O.quat_angle_deg(qa, ref).max() <= 1e-2
    sub = stream_post.decode_ori(logits[1000:1100], is_logits=True)["quat"].cpu().numpy()
    assert O.quat_angle_deg(sub, qa[1000:1100]).max() <= 1e-4
    idx = np.arange(0, B, 607)[:128]
    want_q, _ = O.ori_decode_batch(O.softmax(logits[idx].cpu().numpy()), hist)
    assert O.quat_angle_deg(qa[idx], want_q).max() <= QUAT_TOL_DEG


@pytest.mark.parametrize("n_dim,delete", [(8, False), (6, False), (8, True)])
def test_half_warp_decode_small_histograms(post, stream_post, n_dim, delete):
    """decode_ori_half_kernel (n <= 512 bins, half a warp per image, f32 reduction): taken by itself for quaternion-only calls
    at batches that fill the GPU.  512 bins (full 64-bin slices), 216 bins and the pruned 8-bin histogram (ragged slices), an odd
    batch (the last pair has one image), logits and pdf input, a NaN image; against the oracle on a sample, against the
    warp-per-image kernel on everything, run-to-run determinism and independence of the batch position."""
    hist, _ = O.ori_histogram(n_dim, delete)
    n = hist.shape[0]
    assert n <= 512
    if n % 4:
        pytest.skip("rows of this histogram are not 16-byte multiples: the streaming kernels do not take them")
    post.set_ori_histogram(hist)
    stream_post.set_ori_histogram(hist)
    g = torch.Generator().manual_seed(11 + n_dim)
    B = 16 * torch.cuda.get_device_properties(0).multi_processor_count * 3 + 37
    logits = (torch.randn((B, n), generator=g) * 3).cuda()
    logits[B - 1, 3] = float("nan")
    l0 = stream_post.launch_count()
    a = stream_post.decode_ori(logits, is_logits=True)
    b = stream_post.decode_ori(logits, is_logits=True)
    assert torch.equal(a["quat"][:-1], b["quat"][:-1])
    fl = a["flags"].cpu().numpy()
    assert fl[-1] == 1 and not fl[:-1].any() and torch.isnan(a["quat"][-1]).all()
    qa = a["quat"].cpu().numpy()[:-1]
    ref = np.concatenate([post.decode_ori(logits[i:i + 2048], is_logits=True)["quat"].cpu().numpy() for i in range(0, B - 1, 2048)])[:B - 1]
    assert O.quat_angle_deg(qa, ref).max() <= 1e-2
    sub = stream_post.decode_ori(logits[1001:1101], is_logits=True)["quat"].cpu().numpy()   # small batch: the one-image-per-warp kernel
    assert O.quat_angle_deg(sub, qa[1001:1101]).max() <= 1e-2
    idx = np.arange(0, B - 1, 97)[:96]
    soft = O.softmax(logits[idx].cpu().numpy())
    want_q, _ = O.ori_decode_batch(soft, hist)
    assert O.quat_angle_deg(qa[idx], want_q).max() <= QUAT_TOL_DEG
    # pdf input through the same kernel
    pdf = torch.softmax(logits[:-1], 1)
    qp = stream_post.decode_ori(pdf, is_logits=False)["quat"].cpu().numpy()
    assert O.quat_angle_deg(qp[idx], want_q).max() <= QUAT_TOL_DEG
    assert O.quat_angle_deg(qp, qa).max() <= 1e-2
