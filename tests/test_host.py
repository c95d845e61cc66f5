"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic of the Python mirror,
and the device Jacobi solver compiled for the host.  No compute call needs a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import spef_oracle as O
from spef_b200 import _ffi
from spef_b200.modeling import arch, import_model, copy_state_dict
from spef_b200.spe import OrientationSoftClassification, PositionSoftClassification, SPEUtils
from spef_b200.tools import RunningAverage, synthetic
from spef_b200.tools.evaluation import mad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spef_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spef_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libspef_b200.so does not export {n}"
        assert n in _ffi.SIGNATURES, f"{n} is declared in the header but has no ctypes signature"
    assert set(_ffi.SIGNATURES) == set(names)
    assert lib.spef_abi_version() == 1


def test_config_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "spef_b200.h")).read()
    body = text[text.index("typedef struct spef_config {"):text.index("} spef_config;")]
    fields = re.findall(r"int32_t\s+(\w+);", body)
    assert fields == [f for f, _ in _ffi.SpefConfig._fields_]
    body = text[text.index("typedef struct spef_temporal_out {"):text.index("} spef_temporal_out;")]
    fields = re.findall(r"\*\s*(\w+);", body)
    assert fields == [f for f, _ in _ffi.SpefTemporalOut._fields_]


@pytest.mark.skipif(not NO_GPU, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = _ffi.lib()
    cfg = _ffi.SpefConfig(ctypes.sizeof(_ffi.SpefConfig), 0, 240, 384, 1728, 3, 0, 1, 4, 0)
    h = ctypes.c_void_p()
    rc = lib.spef_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc == 2 and not h.value
    assert b"no CPU fallback" in lib.spef_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        from spef_b200.engine import Engine
        Engine()


def test_create_rejects_bad_config():
    lib = _ffi.lib()
    h = ctypes.c_void_p()
    cfg = _ffi.SpefConfig(4, 0, 240, 384, 1728, 3, 0, 1, 4, 0)  # wrong struct_size
    assert lib.spef_create(ctypes.byref(h), ctypes.byref(cfg)) == 1
    cfg = _ffi.SpefConfig(ctypes.sizeof(_ffi.SpefConfig), 0, 240, 384, 1728, 5, 0, 1, 4, 0)  # regression head with n_pos != 3
    assert lib.spef_create(ctypes.byref(h), ctypes.byref(cfg)) == 1
    assert lib.spef_create(None, None) == 1


def test_device_jacobi_solver_matches_lapack():
    """The 4x4 cyclic Jacobi used by the decode kernel, compiled for the host, against np.linalg.eigh."""
    lib = _ffi.lib()
    rs = np.random.RandomState(0)
    hist, _ = O.ori_histogram(12)
    mats = []
    for sigma in (1e-9, 1.0, 3.0, 10.0):
        p = O.softmax((rs.randn(4, 1728) * sigma).astype(np.float32)).astype(np.float64)
        for i in range(4):
            mats.append(np.einsum("b,bi,bj->ij", p[i], hist, hist))
    mats.append(np.diag([0.1, 0.2, 0.3, 0.4]))
    mats.append(np.outer(hist[5], hist[5]))  # rank 1
    for a in mats:
        ev, evec = np.zeros(4), np.zeros((4, 4))
        a = np.ascontiguousarray(a)
        assert lib.spef_debug_jacobi4_host(a.ctypes.data, ev.ctypes.data, evec.ctypes.data) == 0
        w, v = np.linalg.eigh(a)
        np.testing.assert_allclose(np.sort(ev), w, rtol=1e-12, atol=1e-15)
        q = evec[:, np.argmax(ev)]
        assert O.quat_angle_deg(q / np.linalg.norm(q), v[:, -1]) < 1e-6
        np.testing.assert_allclose(evec @ np.diag(ev) @ evec.T, a, atol=1e-14)


def test_stream_decode_solver_matches_lapack(golden):
    """The eigen-solve of the large-batch decode kernel (f32 Jacobi + one f64 polish step, cofactor inverse), compiled for
    the host, against np.linalg.eig / np.linalg.inv as classification_utils.py:137-142 uses them; gate 0.05 deg."""
    lib = _ffi.lib()
    rs = np.random.RandomState(1)
    hist, _ = O.ori_histogram(12)
    g = golden("encode_decode")
    cases = []
    for sigma in (1e-9, 0.1, 1.0, 3.0, 10.0, 30.0):
        z = (rs.randn(6, 1728) * sigma).astype(np.float32)
        w = np.exp(z.astype(np.float64) - z.max(1, keepdims=True))    # unnormalised weights, S != 1
        cases += [(wi, True) for wi in w]
    cases += [(p.astype(np.float64), False) for p in g["enc_ori"][:16]]  # encoded SPEED labels: sharp pdfs
    worst = 0.0
    for w, is_logits in cases:
        a = np.einsum("b,bi,bj->ij", w, hist, hist)
        sums = np.array([w.sum(), a[0, 0], a[0, 1], a[0, 2], a[0, 3], a[1, 1], a[1, 2], a[1, 3], a[2, 2], a[2, 3], a[3, 3]])
        q, hinv = np.zeros(4, np.float32), np.zeros(16, np.float32)
        assert lib.spef_debug_decode_solve_host(sums.ctypes.data, int(is_logits), q.ctypes.data, hinv.ctypes.data) == 0
        an = a / w.sum() if is_logits else a
        ev, evec = np.linalg.eigh(an)
        gap = (ev[-1] - ev[-2]) / ev[-1]
        ang = float(O.quat_angle_deg(q, evec[:, -1]))
        if gap > 1e-6:   # the dominant direction is defined
            worst = max(worst, ang)
            assert ang < 1e-3, (ang, gap)
        assert abs(np.linalg.norm(q) - 1) < 1e-6 and q[0] >= 0
        want = np.linalg.inv(an)
        assert np.abs(hinv.reshape(4, 4) - want).max() <= 1e-5 * np.abs(want).max()
    bad = np.full(11, np.nan)
    assert lib.spef_debug_decode_solve_host(bad.ctypes.data, 1, q.ctypes.data, None) != 0


def test_resize_tap_tables_match_oracle():
    """The tap tables the library builds for spef_resize_frames (host code, float64) against the oracle's restatement of
    Pillow's coefficient computation -- which tests/test_oracle_golden.py pins to the real torchvision + Pillow transform."""
    import ctypes as C
    lib = _ffi.lib()
    for in_size, out_size in [(1920, 384), (1200, 240), (1200, 1200), (123, 240), (257, 384), (5000, 384), (3000, 240), (1, 7), (997, 384)]:
        b, k = O.resize_coeffs(in_size, out_size)
        first, count = np.zeros(out_size, np.int32), np.zeros(out_size, np.int32)
        coef = np.zeros(out_size * k.shape[1], np.int32)
        ks = C.c_int32(0)
        assert lib.spef_debug_resize_taps_host(in_size, out_size, first.ctypes.data, count.ctypes.data, coef.ctypes.data, coef.size,
                                               C.addressof(ks)) == 0
        assert ks.value == k.shape[1]
        np.testing.assert_array_equal(first, b[:, 0])
        np.testing.assert_array_equal(count, b[:, 1])
        np.testing.assert_array_equal(coef.reshape(out_size, -1), k)


def test_arch_and_state_dict_spec(golden):
    spec = arch.state_dict_spec(1728, 3)
    assert len(spec) == 316 and len(arch.conv_layers()) == 52
    n_param = sum(int(np.prod(s)) for k, s, r in spec if r in ("conv", "bn_weight", "bn_bias", "linear_weight", "linear_bias"))
    assert n_param == 4441283
    assert [b["idx"] for b in arch.block_table() if b["residual"]] == [3, 5, 6, 8, 9, 10, 12, 13, 15, 16]
    assert arch.output_hw(240, 384) == (8, 12)
    # oracle and package agree on the topology
    assert [(b["cin"], b["cout"], b["stride"]) for b in arch.block_table()] == [(b["cin"], b["cout"], b["stride"]) for b in O.block_table()]


def test_synthetic_inputs_are_deterministic():
    a, b = synthetic.synthetic_state_dict(1728, 3), synthetic.synthetic_state_dict(1728, 3)
    assert list(a.keys()) == [k for k, _, _ in arch.state_dict_spec(1728, 3)]
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert float(a["features.features.0.1.running_var"].min()) > 0
    assert torch.equal(synthetic.synthetic_images(2), synthetic.synthetic_images(2))
    t = synthetic.synthetic_targets(64)
    np.testing.assert_allclose(np.linalg.norm(t["ori"], axis=1), 1, atol=1e-6)
    assert t["pos"][:, 2].min() >= 3 and t["pos"][:, 2].max() <= 35


def test_facade_histograms_and_encode_match_reference(golden):
    g, ed = golden("histograms"), golden("encode_decode")
    for n in (8, 12, 16):
        for delete in (False, True):
            o = OrientationSoftClassification(n, 3, delete)
            tag = f"ori{n}_{'del' if delete else 'all'}"
            np.testing.assert_allclose(o.histogram, g[tag + "_hist"], rtol=0, atol=1e-15)
            np.testing.assert_array_equal(o.redundant_flags, g[tag + "_red"])
            assert o.n_bins == g[tag + "_hist"].shape[0]
    o = OrientationSoftClassification(12, 3, False)
    assert o.b.shape == (1728, 4, 4)
    p = PositionSoftClassification(10, 100, np.array([-16, -12, -2]), np.array([16, 12, 40]))
    np.testing.assert_array_equal(p.histogram, g["pos10_hist"])
    for i in range(4):
        np.testing.assert_allclose(o.encode(ed["labels_q"][i]), ed["enc_ori"][i], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(p.encode(ed["labels_t"][i]), ed["enc_pos"][i], rtol=1e-5, atol=1e-12)


def test_speutils_surface():
    su = SPEUtils(None, 'classification', 12, 3, False, 'regression', 10, 100, None)
    assert su.orientation.n_bins == 1728 and su.position.n_bins == 1000 and su.keypoints is None
    with pytest.raises(NotImplementedError):
        SPEUtils(None, 'keypoints', pos_mode='keypoints', keypoints_path='x.mat')
    m = SPEUtils.metrics_from_sums(np.array([2.0, 1.0, 8.0, 4.0, 0, 0, 0, 0]))
    assert m["ori_score"] == np.float32(0.5) and m["pos_score"] == np.float32(0.25) and m["pos_error"] == 2.0
    assert m["esa_score"] == np.float32(0.75) and abs(m["ori_error"] - 0.5 * 180 / np.pi) < 1e-5


def test_import_model_api_on_cpu():
    data = {"valid": [({"torch": torch.rand(2, 3, 240, 384)}, {})]}
    model, bit_width = import_model(data, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification',
                                    n_ori_bins=1728, pos_mode='regression')
    assert bit_width is None and not model.training
    sd = model.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in arch.state_dict_spec(1728, 3)] and len(sd) == 316
    assert hasattr(model, "features") and hasattr(model, "head")
    model.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
    assert torch.equal(model.state_dict()["head.pos.0.bias"], torch.tensor([0.0, 0.0, 10.0]))
    assert set(copy_state_dict(sd, model.state_dict()).keys()) == set(sd.keys())
    assert model.precision == "fp32"  # the reference model family is FP32: BF16 is opt-in (precision='bf16')
    # manual_copy semantics of the reference (model.py:92-119): per key family, by position; foreign keys (a quantised
    # checkpoint's act_quant entries) are ignored unless act_quant=True; a mis-shaped copy is refused
    src = {("module." + k): v + 1 for k, v in synthetic.synthetic_state_dict(1728, 3).items()}
    src["module.features.0.act_quant.scale"] = torch.ones(1)
    out = copy_state_dict(src, dict(model.state_dict()))
    assert list(out.keys()) == list(sd.keys())
    assert torch.equal(out["head.ori.1.weight"], src["module.head.ori.1.weight"])
    assert torch.equal(out["features.features.0.1.running_var"], src["module.features.features.0.1.running_var"])
    bad = dict(src)
    bad["module.head.ori.1.weight"] = torch.zeros(5, 5)
    with pytest.raises(ValueError, match="shape mismatch"):
        copy_state_dict(bad, dict(model.state_dict()))
    with pytest.raises(NotImplementedError):
        import_model(data, 'mobilenet_v2_brevitas', 'ursonet_brevitas', n_ori_bins=1728)
    with pytest.raises(AssertionError):
        import_model(data, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification')  # n_ori_bins missing
    with pytest.raises(NotImplementedError):
        model.train()
    if NO_GPU:  # the forward must fail loudly, never fall back to torch ops
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model(torch.rand(1, 3, 240, 384))


def test_running_average_and_mad():
    ra = RunningAverage(keys=("a",))
    ra.update({"a": 1.0}, 3)
    ra.update({"a": 5.0}, 1)
    assert ra.get("a") == 2.0 and ra.get_multiple(("a",)) == {"a": 2.0}
    assert mad([1.0, 2.0, 3.0, 4.0, 100.0]) == 1.0


def test_host_pack_rounds_like_the_device():
    """spef_pack_bf16_host (the host half of the packed image upload, csrc/host_pack.cpp) == round-to-nearest-even BF16 of every
    float, the stem's own first step (cvt.rn.bf16.f32): ties, carries into the exponent, overflow to inf, denormals kept,
    NaN stays NaN; unaligned heads / ragged tails of the vector code; several pool generations back to back."""
    from spef_b200 import _ffi
    lib = _ffi.lib()
    g = torch.Generator().manual_seed(3)
    special = torch.tensor([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 0.999999, 3.3895314e38, 3.4028235e38, float("inf"), -float("inf"),
                            1e-40, -1e-45, 1.1754944e-38, 0.5 + 2 ** -9, 0.5 + 2 ** -9 + 2 ** -20, float("nan")], dtype=torch.float32)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (300000,), generator=g, dtype=torch.int64).to(torch.int32).view(torch.float32)
    for n, off in ((16, 0), (special.numel(), 0), (1000, 3), (65536 * 3 + 77, 1), (300000, 0), (300000, 5)):
        src = torch.cat([special, torch.rand(400000, generator=g), bits])[off:off + n].contiguous()
        dst = torch.zeros(n + 64, dtype=torch.int16)
        d = dst[off:off + n]
        assert lib.spef_pack_bf16_host(src.data_ptr(), d.data_ptr(), n) == 0
        want = src.to(torch.bfloat16)
        got = d.view(torch.bfloat16)
        nan = torch.isnan(src)
        assert torch.equal(torch.isnan(got), nan)
        assert torch.equal(got[~nan].view(torch.int16), want[~nan].view(torch.int16))
        assert int(dst[:off].abs().sum()) == 0 and int(dst[off + n:].abs().sum()) == 0   # nothing written outside


def test_uint8_pixel_to_bf16_without_the_table():
    """The fused stem turns a uint8 pixel into its BF16 operand as bf16(float(u8) * float(1 / 255)) (csrc/fused_block_t.cuh); the
    reference tensor is float(u8) / 255 (ToTensor), whose BF16 rounding the table-based paths store: equal for all 256 values,
    although the two float32 values differ for about half of them; and the u8 -> float trick (2^23 + u8) - 2^23 is exact."""
    u = np.arange(256, dtype=np.float32)
    ref = torch.from_numpy((u / np.float32(255)).astype(np.float32)).to(torch.bfloat16).view(torch.int16)
    mul = torch.from_numpy((u * np.float32(1.0 / 255.0)).astype(np.float32)).to(torch.bfloat16).view(torch.int16)
    assert torch.equal(ref, mul)
    magic = (np.arange(256, dtype=np.uint32) | np.uint32(0x4B000000)).view(np.float32)
    assert np.array_equal(magic - np.float32(8388608.0), u)
    # the kernel's single FMA: fma(2^23 + u8, c, -(2^23 * c)) -- exact in float64 (24 x 24 bit product), rounded once to float32
    c = np.float32(1.0 / 255.0)
    off = np.float32(-8388608.0) * c
    assert np.float64(off) == -8388608.0 * np.float64(c)                      # 2^23 * c is exact
    fma = (magic.astype(np.float64) * np.float64(c) + np.float64(off)).astype(np.float32)
    assert np.array_equal(fma, (u * c).astype(np.float32))
