"""GPU parity of the Mobile-URSONet forward (through the C ABI) against the CPU oracle and reference goldens.

Gates (BASELINE.json north_star, read per SURVEY.md sections 0.4 / 7.2.2):
  FP32 path           : logits within 1e-4 relative (max|diff| / max|ref|) of the reference's FP32 logits.
  BF16 path, per layer: teacher-forced on the oracle's BF16 activations, every kernel within 1 BF16 ulp of the
                        oracle op with identical rounding points (immune to the random net's chaos).
  BF16 path, end2end  : against the fake-BF16 oracle (same rounding points); the deviation from the FP32 reference
                        is *reported* (an untrained MobileNetV2 amplifies perturbations ~7x per 1e-4).
"""
import numpy as np
import pytest
import torch

from oracle import spef_oracle as O
from spef_b200.tools import synthetic

pytestmark = pytest.mark.gpu

FP32_LOGITS_RTOL = 1e-4


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


def nhwc(x, dtype):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


@pytest.fixture(scope="module")
def sd():
    return synthetic.synthetic_state_dict(1728, 3)


@pytest.fixture(scope="module")
def images():
    return synthetic.synthetic_images(4)


def _engine(sd, precision, pw_impl=0, n_pos=3, max_batch=8):
    from spef_b200.engine import Engine
    eng = Engine(240, 384, 1728, n_pos, n_pos != 3, precision, max_batch, None, pw_impl)
    eng.load_state_dict(sd)
    return eng


def _oracle_layer_io(sd, x, bf16):
    """(input, residual, output) NCHW f32 tensors for the 52 conv layers, chained like the real forward."""
    ios, cur, block_in = [], x, None
    for layer in O.folded_layers(sd):
        if layer["kind"] == "stem" or layer["prefix"].endswith("conv.0") or layer["prefix"].endswith(".18"):
            block_in = cur
        res = block_in if layer["residual"] else None
        out = O.apply_layer(layer, cur, res, bf16)
        ios.append((cur, res, out))
        cur = out
    return ios


def _spread(t, batch, pick):
    """[len(pick), ...] -> [batch, ...]: the picked slots hold the oracle's images, every other slot a copy of the first."""
    if batch == len(pick):
        return t
    full = t[:1].repeat(batch, *([1] * (t.dim() - 1)))
    full[pick] = t
    return full


def _check_every_layer(sd, images, precision, pw_impl, batch, pick):
    """Each of the 54 launches of the forward on the oracle's input for that layer.  batch / pick: the launch runs `batch` images,
    the oracle checks the slots in `pick` (images are independent, so three slots of a 256-image launch -- first, middle, last --
    cover what changes with the batch: persistent CTAs looping over many tiles, ring slots wrapping, tile tails)."""
    bf16 = precision == "bf16"
    eng = _engine(sd, precision, pw_impl, max_batch=max(8, batch))
    dt = torch.bfloat16 if bf16 else torch.float32
    x = images[:len(pick)]
    ios = _oracle_layer_io(sd, x, bf16)
    assert eng.num_layers() == len(ios) + 2
    worst = 0.0
    for i, (inp, res, want) in enumerate(ios):
        info = eng.layer_info(i)
        got = eng.layer_forward(i, _spread(inp if i == 0 else nhwc(inp, dt), batch, pick), None if res is None else _spread(nhwc(res, dt), batch, pick))
        got = got[pick].float().cpu()
        want_nhwc = want.permute(0, 2, 3, 1).contiguous()
        assert got.shape == want_nhwc.shape, (i, info)
        scale = float(want_nhwc.abs().max())
        err = (got - want_nhwc).abs()
        if bf16:
            # identical rounding points: differences are FP32 accumulation-order effects that flip a BF16 rounding
            ulp = torch.maximum(want_nhwc.abs(), torch.tensor(scale * 2 ** -8)) * 2 ** -7
            frac_off = float((err > 0).float().mean())
            assert float((err / ulp).max()) <= 1.01, f"layer {i} {info}: > 1 BF16 ulp"
            assert frac_off < 0.02, f"layer {i} {info}: {frac_off:.4f} of the elements differ"
        else:
            assert float(err.max()) <= 2e-5 * scale, f"layer {i} {info}: rel err {float(err.max()) / scale:.2e}"
        worst = max(worst, float(err.max()) / scale)
    # global mean (layer 52) and head GEMM (layer 53)
    last = ios[-1][2]
    feat = last.permute(0, 2, 3, 1).reshape(len(pick), -1, 1280).contiguous()
    pooled = eng.layer_forward(52, _spread(feat.to(dt), batch, pick))[pick].float().cpu()
    want_pool = last.mean([2, 3])
    want_pool = O.bf16_round(want_pool) if bf16 else want_pool
    assert float((pooled - want_pool).abs().max()) <= (2 ** -7 if bf16 else 1e-5) * float(want_pool.abs().max())
    wo, wp = sd["head.ori.1.weight"], sd["head.pos.0.weight"]
    if bf16:
        wo, wp = O.bf16_round(wo), O.bf16_round(wp)
    head = eng.layer_forward(53, _spread(want_pool.to(dt), batch, pick))[pick].cpu()
    want_head = torch.cat([torch.nn.functional.linear(want_pool, wo, sd["head.ori.1.bias"]),
                           torch.nn.functional.linear(want_pool, wp, sd["head.pos.0.bias"])], dim=1)
    assert head.shape[1] == eng.layer_info(53)["cout"] >= want_head.shape[1]
    assert rel(head[:, :want_head.shape[1]].numpy(), want_head.numpy()) < 2e-5
    assert float(head[:, want_head.shape[1]:].abs().max() if head.shape[1] > want_head.shape[1] else 0.0) == 0.0
    print(f"[{precision} pw_impl={pw_impl} B={batch}] worst per-layer relative error {worst:.3e}")


@pytest.mark.parametrize("precision,pw_impl", [("fp32", 0), ("bf16", 1), ("bf16", 0)])
def test_every_layer_teacher_forced(sd, images, precision, pw_impl):
    _check_every_layer(sd, images, precision, pw_impl, 2, [0, 1])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_every_layer_teacher_forced_full_batch(sd, images, precision):
    """The same gates inside a launch of the benchmark shape (B = 256): slots 0, 127 and 255 against the oracle."""
    _check_every_layer(sd, images, precision, 0, 256, [0, 127, 255])


def _oracle_block(layers, first, last, x, hidden_fp32):
    """Oracle output of one InvertedResidual block (layers first..last) on NCHW input x with the rounding points of its kernel:
    the channel-lane fused kernel keeps the hidden tensor (expand output) in FP32, everything else rounds every layer output."""
    cur = x
    for li in range(first, last + 1):
        layer = layers[li]
        cur = O.apply_layer(layer, cur, x if layer["residual"] else None, True,
                            round_out=not (hidden_fp32 and layer["role"] == "expand"))
    return cur


def _check_fused_blocks(sd, images, batch=3, pick=None):
    """Every fused InvertedResidual kernel (expand -> depthwise -> project [+ x] in one launch; for the wide blocks the expand
    GEMM followed by ONE depthwise -> project kernel, variant 3) on the oracle's BF16 input of that block.
      * against the oracle WITH THE KERNEL'S ROUNDING POINTS (the staged kernel rounds the hidden tensor to BF16 like the
        per-layer kernels, the channel-lane kernel keeps it in FP32): the depthwise outputs are not teacher-forced inside a block,
        so a 1-ulp flip of a BF16 depthwise value (FP32 accumulation order) moves a few block outputs by a few output ulps:
        < 1 % of the elements may differ and none by more than 32 BF16 ulp floored at 2^-8 of the tensor scale (= 1e-3 of the scale:
        one flipped depthwise value times its project weight, seen on an output that happens to cancel);
      * against the chain of per-layer kernels (each within 1 BF16 ulp of the oracle, test above): the staged kernel has the
        same rounding points and FP32 accumulation order and is bit-identical, and so is the depthwise -> project kernel (the depthwise
        output is rounded once to BF16 either way; it just never leaves the SM); the channel-lane kernel differs by the hidden
        tensor's BF16 rounding (2^-9 relative per hidden element), which is reported and bounded at 16 output ulp.
    batch / pick: the images of the batch the oracle checks (the B = 256 test teacher-forces images 0, 127, 255 of a full batch)."""
    eng = _engine(sd, "bf16", 0, max_batch=max(8, batch))
    pick = list(range(batch)) if pick is None else pick
    x = images[:len(pick)]
    ios = _oracle_layer_io(sd, x, True)
    layers = O.folded_layers(sd)
    n_fused = {1: 0, 2: 0, 3: 0}
    for bi in range(eng.num_blocks()):
        info = eng.block_info(bi)
        if not info["fused"]:
            with pytest.raises(Exception, match="no fused kernel"):
                eng.block_forward(bi, nhwc(ios[info["first_layer"]][0], torch.bfloat16))
            continue
        n_fused[info["fused"]] += 1
        first, last = info["first_layer"], info["first_layer"] + info["n_layers"] - 1
        inp = _spread(nhwc(ios[first][0], torch.bfloat16), batch, pick)
        want = _oracle_block(layers, first, last, ios[first][0], hidden_fp32=(info["fused"] == 2)).permute(0, 2, 3, 1).contiguous()
        got_full = eng.block_forward(bi, inp)
        got = got_full[pick].float().cpu()
        assert got.shape == want.shape, (bi, info)
        cur = inp
        for li in range(first, last + 1):
            cur = eng.layer_forward(li, cur, inp if eng.layer_info(li)["residual"] else None)
        chain = cur[pick].float().cpu()
        scale = float(want.abs().max())
        if info["fused"] in (1, 3):
            np.testing.assert_array_equal(got.numpy(), chain.numpy(), err_msg=f"block {bi} {info} vs per-layer kernels")
        else:
            err = (got - chain).abs()
            ulp = torch.maximum(chain.abs(), torch.tensor(scale * 2 ** -8)) * 2 ** -7
            # the per-layer chain rounds the hidden tensor to BF16 (2^-9 relative per hidden element): the block outputs differ by a
            # few 1e-3 of the summands' magnitude -- many ulps of an output that happens to cancel -- so this bound is relative to
            # the tensor scale; the tight gate is the oracle with the kernel's own rounding points below
            print(f"block {bi}: channel-lane vs per-layer kernels (hidden tensor FP32 vs BF16): max |diff| / scale {float(err.max()) / scale:.2e}, "
                  f"max {float((err / ulp).max()):.1f} floored ulp, {float((err > 0).float().mean()):.3f} of the elements differ")
            assert float(err.max()) <= 1e-2 * scale, f"block {bi} {info} vs per-layer kernels"
        err = (got - want).abs()
        ulp = torch.maximum(want.abs(), torch.tensor(scale * 2 ** -8)) * 2 ** -7
        worst, frac_off = float((err / ulp).max()), float((err > 0).float().mean())
        print(f"block {bi}: vs oracle max {worst:.2f} ulp, {frac_off:.4f} of the elements differ")
        assert worst <= 32.0, f"block {bi} {info} vs oracle: {worst:.2f} BF16 ulp"
        assert frac_off < 0.01, f"block {bi} {info} vs oracle: {frac_off:.4f} of the elements differ"
        if batch != len(pick):   # every other slot holds a copy of image 0: bit-identical to slot pick[0]
            ref0 = got_full[pick[0]]
            assert all(torch.equal(got_full[j], ref0) for j in range(batch) if j not in pick), f"block {bi}: batch slots differ"
    return n_fused


def test_fused_blocks_teacher_forced(sd, images):
    n = _check_fused_blocks(sd, images)
    assert n[1] + n[2] >= 8, f"only {n} blocks are fused"
    assert n[1] + n[2] + n[3] >= 16, f"only {n} of the 17 blocks run fused kernels"   # all but the stride-2 block 14 (576 -> 160)


def _check_stem_block(sd, images, batch=3, pick=None, u8=False):
    """Stem conv + first InvertedResidual block as ONE kernel (mobilenet_v2.py:252-262), teacher-forced on the image:
      * against the oracle with the kernel's rounding points -- image taps and stem weights BF16, the stem output (the kernel's hidden tensor)
        FP32, depthwise output BF16, block output BF16 -- at the gates of the fused blocks (32 floored BF16 ulp, < 1 % of the elements);
      * against the separate stem launch followed by the block's own fused kernel (which rounds the stem output to BF16): reported, bounded.
    uint8 images: the same pixels as bytes must give bit-identical results to their float32 values / 255."""
    eng = _engine(sd, "bf16", 0, max_batch=max(8, batch))
    assert eng.stem_fusion_active(), "the stem is not fused into the first block"
    assert 0 in eng.fp32_hidden_blocks()
    pick = list(range(batch)) if pick is None else pick
    x = images[:len(pick)]
    if u8:
        xb = (x * 255.0).round().clamp(0, 255).to(torch.uint8)
        x = xb.float() / 255.0
    layers = O.folded_layers(sd)
    info = eng.block_info(0)
    first, last = info["first_layer"], info["first_layer"] + info["n_layers"] - 1
    assert first == 1
    stem_fp32 = O.apply_layer(layers[0], x, None, True, round_out=False)
    want = _oracle_block(layers, first, last, stem_fp32, hidden_fp32=True).permute(0, 2, 3, 1).contiguous()
    full = _spread(x, batch, pick)
    if u8:
        eng.set_image_dtype(torch.uint8)
        got_full = eng.stem_block_forward(_spread(xb, batch, pick))
        eng.set_image_dtype(torch.float32)
        got_f32 = eng.stem_block_forward(full)
        assert torch.equal(got_full, got_f32), "uint8 and float32 images differ through the fused stem"
    else:
        got_full = eng.stem_block_forward(full)
    got = got_full[pick].float().cpu()
    assert got.shape == want.shape
    eng.set_stem_fusion(False)
    assert not eng.stem_fusion_active() and 0 not in eng.fp32_hidden_blocks()
    stem_out = eng.layer_forward(0, full.to(eng.device), None)
    chain = eng.block_forward(0, stem_out)[pick].float().cpu()
    eng.set_stem_fusion(True)
    scale = float(want.abs().max())
    err = (got - chain).abs()
    print(f"stem + block 0: one kernel vs stem launch + block kernel (stem output FP32 vs BF16): max |diff| / scale {float(err.max()) / scale:.2e}, "
          f"{float((err > 0).float().mean()):.3f} of the elements differ")
    assert float(err.max()) <= 1e-2 * scale
    err = (got - want).abs()
    ulp = torch.maximum(want.abs(), torch.tensor(scale * 2 ** -8)) * 2 ** -7
    worst, frac_off = float((err / ulp).max()), float((err > 0).float().mean())
    print(f"stem + block 0: vs oracle max {worst:.2f} ulp, {frac_off:.4f} of the elements differ")
    assert worst <= 32.0, f"stem + block 0 vs oracle: {worst:.2f} BF16 ulp"
    assert frac_off < 0.01, f"stem + block 0 vs oracle: {frac_off:.4f} of the elements differ"
    if batch != len(pick):
        ref0 = got_full[pick[0]]
        assert all(torch.equal(got_full[j], ref0) for j in range(batch) if j not in pick), "stem + block 0: batch slots differ"


@pytest.mark.parametrize("u8", [False, True])
def test_stem_block_teacher_forced(sd, images, u8):
    _check_stem_block(sd, images, u8=u8)


def test_stem_block_teacher_forced_full_batch(sd, images):
    """Inside a B = 256 launch: slots 0, 127, 255 against the oracle, all other slots bit-identical copies."""
    _check_stem_block(sd, images, batch=256, pick=[0, 127, 255])


def test_stem_fusion_forward_equals_separate_stem(sd):
    """End to end: logits with the stem inside the first block's kernel vs the separate stem launch (ragged batch)."""
    eng = _engine(sd, "bf16", 0)
    x = synthetic.synthetic_images(5, seed=11)
    o_f, p_f = [t.cpu().numpy() for t in eng.forward(x)]
    eng.set_stem_fusion(False)
    o_l, p_l = [t.cpu().numpy() for t in eng.forward(x)]
    eng.set_stem_fusion(True)
    print(f"fused stem vs separate stem: logits {rel(o_f, o_l):.2e}, pos {rel(p_f, p_l):.2e}")
    assert rel(o_f, o_l) < 3e-2 and rel(p_f, p_l) < 3e-2
    assert (o_f.argmax(1) == o_l.argmax(1)).mean() >= 0.8


def test_dw_project_kernel_on_every_shape_it_takes(sd, images, monkeypatch):
    """SPEF_DWP_FORCE=1 runs the expand GEMM + depthwise -> project kernel on every stride-1 block it has a plan for (hidden widths
    192 / 384 / 576 / 960, maps 30x48 ... 8x12, with and without skip connection, one- and two-half accumulators): bit-identical
    to the per-layer kernels, within the block gates of the oracle."""
    monkeypatch.setenv("SPEF_DWP_FORCE", "1")
    n = _check_fused_blocks(sd, images)
    assert n[3] >= 10, f"only {n} blocks took the depthwise -> project kernel"


def test_fused_blocks_teacher_forced_full_batch(sd, images):
    """Every fused block inside a B = 256 launch: slots 0, 127, 255 against the oracle, all other slots bit-identical copies."""
    n = _check_fused_blocks(sd, images, batch=256, pick=[0, 127, 255])
    assert n[1] + n[2] == 11 and n[3] == 6, f"only {n} of the 17 blocks are fused"   # block 14 (stride 2) takes the depthwise -> project kernel too


def test_fused_block_variants(sd, images, monkeypatch):
    """Default plan: channel-lane kernels for the blocks with Cin <= 64.  SPEF_FB_VARIANT=0 selects the staged kernel for the
    same blocks (bit-identical to the per-layer kernels); SPEF_FBT_NG=2 the two-group channel-lane kernel."""
    assert _check_fused_blocks(sd, images)[2] >= 8
    monkeypatch.setenv("SPEF_FB_VARIANT", "0")
    n = _check_fused_blocks(sd, images)
    assert n[1] >= 8 and n[2] == 0
    monkeypatch.delenv("SPEF_FB_VARIANT")
    monkeypatch.setenv("SPEF_FBT_NG", "2")
    assert _check_fused_blocks(sd, images)[2] >= 8


def test_fused_forward_equals_per_layer_forward(sd):
    """End to end: logits of the fused path vs the per-layer path (same rounding points) on a ragged batch."""
    eng = _engine(sd, "bf16", 0)
    x = synthetic.synthetic_images(5, seed=11)
    o_f, p_f = [t.cpu().numpy() for t in eng.forward(x)]
    eng.set_fusion(False)
    assert not any(eng.block_info(i)["fused"] for i in range(eng.num_blocks()))
    o_l, p_l = [t.cpu().numpy() for t in eng.forward(x)]
    eng.set_fusion(True)
    print(f"fused vs per-layer: logits {rel(o_f, o_l):.2e}, pos {rel(p_f, p_l):.2e}")
    assert rel(o_f, o_l) < 3e-2 and rel(p_f, p_l) < 3e-2
    assert (o_f.argmax(1) == o_l.argmax(1)).mean() >= 0.8


@pytest.mark.parametrize("tag,n_pos", [("murso", 3), ("mursop", 1000)])
def test_fp32_logits_vs_reference_golden(golden, images, tag, n_pos):
    g = golden("network")
    eng = _engine(synthetic.synthetic_state_dict(1728, n_pos), "fp32", n_pos=n_pos)
    ori, pos = eng.forward(images)
    r_ori, r_pos = rel(ori.cpu().numpy(), g[tag + "_ori"]), rel(pos.cpu().numpy(), g[tag + "_pos"])
    print(f"FP32 {tag}: logits rel {r_ori:.2e}, pos rel {r_pos:.2e}")
    assert r_ori <= FP32_LOGITS_RTOL and r_pos <= FP32_LOGITS_RTOL
    np.testing.assert_array_equal(ori.argmax(1).cpu().numpy(), g[tag + "_ori"].argmax(1))  # bin index bit-exact


def test_bf16_end_to_end(golden, sd, images):
    g = golden("network")
    tc = _engine(sd, "bf16", 0)
    simt = _engine(sd, "bf16", 1)
    want_ori, want_pos = O.forward_folded(sd, images, bf16=True, fp32_hidden_blocks=tc.fp32_hidden_blocks())
    o_tc, p_tc = [t.cpu().numpy() for t in tc.forward(images)]
    o_si, p_si = [t.cpu().numpy() for t in simt.forward(images)]
    print(f"BF16 tcgen05 vs fake-BF16 oracle: logits {rel(o_tc, want_ori.numpy()):.2e} pos {rel(p_tc, want_pos.numpy()):.2e}; "
          f"tcgen05 vs SIMT: {rel(o_tc, o_si):.2e}; vs FP32 reference (reported): {rel(o_tc, g['murso_ori']):.2e}")
    assert rel(o_tc, want_ori.numpy()) < 3e-2 and rel(p_tc, want_pos.numpy()) < 3e-2
    assert rel(o_si, want_ori.numpy()) < 3e-2
    assert rel(o_tc, o_si) < 3e-2
    assert rel(o_tc, g["murso_ori"]) < 0.15  # reported bound, not the parity gate
    assert (o_tc.argmax(1) == want_ori.numpy().argmax(1)).mean() >= 0.75


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batch_independence_and_tails(sd, precision):
    """Ragged batches (M tails of every GEMM tile): image i of a batch of 5 equals the same image run alone."""
    eng = _engine(sd, precision)
    x = synthetic.synthetic_images(5, seed=77)
    o5, p5 = eng.forward(x)
    for i in (0, 4):
        o1, p1 = eng.forward(x[i:i + 1])
        np.testing.assert_array_equal(o1.cpu().numpy(), o5[i:i + 1].cpu().numpy())
        np.testing.assert_array_equal(p1.cpu().numpy(), p5[i:i + 1].cpu().numpy())
    o3, _ = eng.forward(x[1:4])
    np.testing.assert_array_equal(o3.cpu().numpy(), o5[1:4].cpu().numpy())
    with pytest.raises(ValueError):
        eng.forward(synthetic.synthetic_images(9))  # > max_batch
    with pytest.raises(ValueError):
        eng.forward(torch.rand(1, 3, 128, 128))


def test_predict_plugin_and_evaluation_dropin(golden):
    """SPEB200.predict has SPETorch's contract; evaluation() reproduces the reference's rec_score / rec_error
    (FP32 engine; reference values from tests/golden/evaluation.npz)."""
    from spef_b200.modeling import import_model
    from spef_b200.spe import SPEB200, SPEUtils
    from spef_b200.tools import evaluation
    g = golden("evaluation")
    su = SPEUtils(None, 'classification', 12, 3, False, 'regression', 10, 100, None)
    loader = synthetic.SyntheticLoader(10, 4)
    model, bw = import_model({"valid": loader}, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification',
                             n_ori_bins=su.orientation.n_bins, pos_mode='regression', precision="fp32")
    model.load_state_dict(synthetic.synthetic_state_dict(1728, 3))
    spe = SPEB200(model, torch.device("cuda:0"), su)
    first = loader.batches[0][0]["torch"]
    pose, ms = spe.predict(first)
    assert set(pose) == {"ori_soft", "ori", "pos"} and ms > 0
    assert pose["ori_soft"].shape == (4, 1728) and pose["ori"].shape == (4, 4) and pose["pos"].shape == (4, 3)
    assert all(v.dtype == np.float32 for v in pose.values())
    assert O.quat_angle_deg(pose["ori"], g["pred_ori"][:4]).max() < 0.05
    np.testing.assert_allclose(pose["pos"], g["pred_pos"][:4], rtol=1e-4)
    np.testing.assert_allclose(pose["ori_soft"].sum(1), 1.0, atol=1e-5)
    pose_dev, _ = spe.predict(first.cuda())  # CUDA tensor input takes the device-pointer entry point
    np.testing.assert_array_equal(pose_dev["ori"], pose["ori"])
    rec_score, rec_error = evaluation(spe, {"valid": loader}, su, ("valid",))
    assert set(rec_score["valid"]) == {"ori", "pos", "esa"} and len(rec_score["valid"]["esa"]) == 1
    np.testing.assert_allclose([rec_score["valid"][k][0] for k in ("ori", "pos", "esa")], g["score"], rtol=1e-4)
    np.testing.assert_allclose([rec_error["valid"][k][0] for k in ("ori", "pos", "ori_std", "pos_std", "ori_mad", "pos_mad")],
                               g["error"], rtol=1e-3)
    # the model is also a legal `model` for the reference's own SPETorch: callable, .to(), .eval(), returns (ori, pos)
    ori, pos = model.to(torch.device("cuda:0")).eval()(first.to("cuda:0"))
    assert ori.shape == (4, 1728) and pos.shape == (4, 3) and ori.is_cuda
    # a generic duck-typed back-end goes through the per-batch route and gives the same numbers
    class Duck:
        def predict(self, images):
            return spe.predict(images)
    rs2, re2 = evaluation(Duck(), {"valid": loader}, su, ("valid",))
    np.testing.assert_allclose(rs2["valid"]["esa"], rec_score["valid"]["esa"], rtol=1e-6)
    np.testing.assert_allclose(re2["valid"]["ori_mad"], rec_error["valid"]["ori_mad"], rtol=1e-6)


def test_inference_dropin_vs_oracle():
    """temporal.Inference on Mobile-URSONet+ (1728 + 1000 heads): per-frame outputs equal the oracle's temporal
    flow applied to the same logits (teacher-forced on the engine's own logits)."""
    from spef_b200.modeling import import_model
    from spef_b200.spe import SPEUtils
    from spef_b200.temporal import Inference
    su = SPEUtils(None, 'classification', 12, 3, False, 'classification', 10, 100, None)
    frames = synthetic.synthetic_images(5, seed=31)
    data = {"v": [({"torch": frames[:1]}, {})]}
    model, _ = import_model(data, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification', n_ori_bins=1728,
                            pos_mode='classification', n_pos_bins=1000, precision="bf16")
    model.load_state_dict(synthetic.synthetic_state_dict(1728, 1000))
    inf = Inference(model, 'gpu_host', su)
    ref = O.TemporalInference(O.ori_histogram(12)[0], O.pos_histogram(10))
    inf.reset()
    for k in range(5):
        still, ms, video = inf.predict(frames[k:k + 1], 'Adaptative')
        lo, lp = inf.engine.forward(frames[k:k + 1])
        rs, rv = ref.step(lo[0].cpu().numpy(), lp[0].cpu().numpy())
        assert still["ori"].shape == (4,) and video["ori_soft"].shape == (1728,) and video["pos_soft"].shape == (1000,)
        assert O.quat_angle_deg(still["ori"], rs["ori"]) <= 0.05 and O.quat_angle_deg(video["ori"], rv["ori"]) <= 0.05
        np.testing.assert_allclose(still["pos"], rs["pos"], rtol=1e-4)
        np.testing.assert_allclose(video["pos"], rv["pos"], rtol=1e-4)
        np.testing.assert_allclose(video["ori_distance"], rv["ori_distance"], rtol=1e-3, atol=1e-7)
    still, _, video = inf.predict(frames[:1])  # video_type None: still pose only
    assert video is None and set(still) == {"ori_soft", "pos_soft", "ori", "pos"}
    with pytest.raises(NotImplementedError):
        Inference(model, 'cpu_host', su)


def test_full_batch_256_properties(sd):
    """BASELINE config 2 size (BF16, B = 256): determinism, batch-slot independence, unit quaternions."""
    from spef_b200.engine import Engine
    eng = Engine(240, 384, 1728, 3, False, "bf16", 256)
    eng.load_state_dict(sd)
    eng.set_ori_histogram(O.ori_histogram(12)[0])
    base = synthetic.synthetic_images(128, seed=5)
    x = torch.cat([base, base]).cuda()
    a = eng.predict(x, want_soft=False, want_argmax=True)
    b = eng.predict(x, want_soft=False, want_argmax=True)
    qa = a["ori"].cpu().numpy()
    np.testing.assert_array_equal(qa, b["ori"].cpu().numpy())            # run-to-run deterministic
    np.testing.assert_array_equal(qa[:128], qa[128:])                     # slot i == slot i + 128
    np.testing.assert_array_equal(a["argmax"].cpu().numpy()[:128], a["argmax"].cpu().numpy()[128:])
    assert np.abs(np.linalg.norm(qa, axis=1) - 1).max() < 1e-6 and not a["flags"].any()
    # 4 of the 256 against the fake-BF16 oracle (same rounding points)
    want_ori, want_pos = O.forward_folded(sd, base[:4], bf16=True, fp32_hidden_blocks=eng.fp32_hidden_blocks())
    o, p = eng.forward(x[:4])
    assert rel(o.cpu().numpy(), want_ori.numpy()) < 3e-2
    want_q, _ = O.ori_decode_batch(O.softmax(o.cpu().numpy()), O.ori_histogram(12)[0])
    assert O.quat_angle_deg(qa[:4], want_q).max() <= 0.05   # decode on identical logits: the 0.05 deg gate


def test_uint8_images_equal_float_images(sd):
    """Input side of the path: uint8 pixels (what ToTensor divides by 255) give bit-identical results to the float
    tensor of the reference contract, because the stem maps a pixel to bf16(float(u8) / 255.0f) itself."""
    eng = _engine(sd, "bf16")
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (3, 3, 240, 384), generator=g, dtype=torch.uint8)
    f32 = u8.float().div(255)          # torchvision ToTensor semantics
    o_f, p_f = eng.forward(f32)
    eng.set_image_dtype(torch.uint8)
    o_u, p_u = eng.forward(u8)
    eng.set_image_dtype(torch.float32)
    np.testing.assert_array_equal(o_u.cpu().numpy(), o_f.cpu().numpy())
    np.testing.assert_array_equal(p_u.cpu().numpy(), p_f.cpu().numpy())
    fp32 = _engine(sd, "fp32")
    with pytest.raises(Exception, match="BF16"):
        fp32.set_image_dtype(torch.uint8)


def test_uint8_host_paths_and_decode_guards(sd):
    """(1) after set_image_dtype(uint8) the host-buffer entry points must be fed uint8 pixels -- predict_host / eval_batch on
    CPU tensors convert (or pass through) accordingly and agree bit for bit with the float route; (2) the fused evaluation route
    raises the reference's ValueError when a decode guard fires (classification_utils.py:134), here for an image full of NaN."""
    from spef_b200.modeling import import_model
    from spef_b200.spe import SPEB200, SPEUtils
    from spef_b200.tools import evaluation
    eng = _engine(sd, "bf16")
    eng.set_ori_histogram(O.ori_histogram(12)[0])
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (3, 3, 240, 384), generator=g, dtype=torch.uint8)
    f32 = u8.float().div(255)
    tg = synthetic.synthetic_targets(3)
    want = eng.predict_host(f32, want_soft=False)
    eng.eval_reset()
    per_f = eng.eval_batch(f32, tg["ori"], tg["pos"], want_per_image=True)
    sums_f = eng.eval_read()
    eng.set_image_dtype(torch.uint8)
    for x in (u8, f32):   # uint8 pixels pass through, a float batch is converted back to pixels
        got = eng.predict_host(x, want_soft=False)
        np.testing.assert_array_equal(got["ori"], want["ori"])
        np.testing.assert_array_equal(got["pos"], want["pos"])
    eng.eval_reset()
    per_u = eng.eval_batch(u8, tg["ori"], tg["pos"], want_per_image=True)
    np.testing.assert_array_equal(per_u, per_f)
    np.testing.assert_array_equal(eng.eval_read(), sums_f)
    eng.set_image_dtype(torch.float32)

    su = SPEUtils(None, 'classification', 12, 3, False, 'regression', 10, 100, None)
    loader = synthetic.SyntheticLoader(8, 4)
    model, _ = import_model({"valid": loader}, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification',
                            n_ori_bins=1728, pos_mode='regression', precision="bf16")
    model.load_state_dict(sd)
    spe = SPEB200(model, torch.device("cuda:0"), su)
    rs, _ = evaluation(spe, {"valid": loader}, su, ("valid",))
    # the same loader as uint8 pixels: the uint8 ingest route is taken automatically and restores the engine's dtype
    class U8Loader:
        def __iter__(self):
            for images, targets in loader:
                yield {"torch": (images["torch"] * 255).round().to(torch.uint8)}, targets
    rs8, _ = evaluation(spe, {"valid": U8Loader()}, su, ("valid",))
    assert spe.engine.image_dtype == torch.float32
    class QuantLoader:   # float images of exactly those pixels: must give the same numbers as the uint8 route
        def __iter__(self):
            for images, targets in loader:
                yield {"torch": (images["torch"] * 255).round().div(255)}, targets
    rsq, _ = evaluation(spe, {"valid": QuantLoader()}, su, ("valid",))
    assert rs8["valid"]["esa"] == rsq["valid"]["esa"] and abs(rs8["valid"]["esa"][0] - rs["valid"]["esa"][0]) < 0.5
    class NanLoader:
        def __iter__(self):
            for images, targets in loader:
                x = images["torch"].clone()
                x[1] = float("nan")
                yield {"torch": x}, targets
    with pytest.raises(ValueError, match="orientation decoding"):
        evaluation(spe, {"valid": NanLoader()}, su, ("valid",))


def test_eval_step_graph_replay_equals_direct_launches(sd, images, monkeypatch):
    """spef_eval_batch replays its ~35 launches as ONE CUDA graph once a call signature has come back (same device buffers, same
    batch): per-image errors bit-identical to direct launches (SPEF_EVAL_GRAPH=0), accumulated sums equal, both host- and
    device-buffer routes; changing the histogram afterwards drops the graphs (they hold table pointers)."""
    tg = synthetic.synthetic_targets(4)
    qt, tt = torch.as_tensor(tg["ori"]).float(), torch.as_tensor(tg["pos"]).float()
    hist = O.ori_histogram(12)[0]
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SPEF_EVAL_GRAPH", mode)
        eng = _engine(sd, "bf16")
        eng.set_ori_histogram(hist)
        x, q, t = images.cuda(), qt.cuda(), tt.cuda()
        eng.eval_reset()
        l0 = eng.launch_count()
        pers = [eng.eval_batch(x, q, t, want_per_image=False) for _ in range(2)]      # direct, direct (signature seen) ...
        per_dev = [eng._empty(4, 2) for _ in range(1)][0]
        from spef_b200._ffi import ptr
        for _ in range(4):                                                               # ... seen, captured + replayed, replayed, replayed
            eng._ck(eng.lib.spef_eval_batch(eng._h, ptr(x), ptr(q), ptr(t), 4, ptr(per_dev), None))
        torch.cuda.synchronize()
        n_launch = eng.launch_count() - l0
        host = [eng.eval_batch(images, qt, tt, want_per_image=True) for _ in range(4)]  # host route: the ctx's own device buffers
        sums = eng.eval_read()
        res[mode] = (per_dev.cpu().numpy().copy(), [h.copy() for h in host], sums, n_launch)
        if mode == "1":
            eng.set_ori_histogram(O.ori_histogram(12)[0])                                # drops the graphs; the next calls run and re-capture
            again = [eng.eval_batch(images, qt, tt, want_per_image=True) for _ in range(3)]
            for a in again:
                np.testing.assert_array_equal(a, host[0])
        eng.close()
    np.testing.assert_array_equal(res["0"][0], res["1"][0])
    for a, b in zip(res["0"][1], res["1"][1]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(res["1"][1][0], res["1"][1][3])
    np.testing.assert_allclose(res["0"][2], res["1"][2], rtol=1e-12)
    assert res["0"][2][3] == 40 and res["0"][3] == res["1"][3], (res["0"][2], res["0"][3], res["1"][3])   # 10 steps of 4 images; launch_count counts replayed launches


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_evaluation_lanes_equal_single_stream(precision):
    """evaluation() issues the batches of a phase round-robin over SPEB200.lanes device contexts / CUDA streams (Engine.lanes):
    same rec_score / rec_error as one lane (the per-context float64 sums are added: linear), per-image statistics from the same
    per-image rows; host batches, device-resident batches and uint8 host batches; a weight reload reaches the twin context."""
    from spef_b200.modeling import import_model
    from spef_b200.spe import SPEB200, SPEUtils
    from spef_b200.tools import evaluation
    su = SPEUtils(None, 'classification', 12, 3, False, 'regression', 10, 100, None)
    loader = synthetic.SyntheticLoader(26, 4)        # 7 batches (the last one ragged): the lanes get 4 and 3

    class DevLoader:
        def __iter__(self):
            for im, tg in loader:
                yield {"torch": im["torch"].cuda()}, tg

    class U8Loader:
        def __iter__(self):
            for im, tg in loader:
                yield {"torch": (im["torch"] * 255).round().to(torch.uint8)}, tg

    def run(lanes, sd):
        model, _ = import_model({"valid": loader}, 'mobilenet_v2_pytorch', 'ursonet_pytorch', ori_mode='classification',
                                n_ori_bins=su.orientation.n_bins, pos_mode='regression', precision=precision)
        model.load_state_dict(sd)
        spe = SPEB200(model, torch.device("cuda:0"), su, lanes=lanes)
        out = [evaluation(spe, {"valid": ld}, su, ("valid",)) for ld in ([loader, DevLoader()] + ([U8Loader()] if precision == "bf16" else []))]
        if lanes > 1:
            assert len(spe.engine._twins) == lanes - 1
            sd2 = {k: (v * 1.01 if k == "head.pos.0.bias" else v) for k, v in sd.items()}
            model.load_state_dict(sd2)
            spe.update_model(model, torch.device("cuda:0"))
            out.append(evaluation(spe, {"valid": loader}, su, ("valid",)))
        spe.delete_model()
        return out
    sd = synthetic.synthetic_state_dict(1728, 3)
    one, two = run(1, sd), run(2, sd)
    for (s1, e1), (s2, e2) in zip(one, two):
        for k in ("ori", "pos", "esa"):
            np.testing.assert_allclose(s2["valid"][k], s1["valid"][k], rtol=1e-12)
        for k in ("ori", "pos", "ori_std", "pos_std", "ori_mad", "pos_mad"):
            np.testing.assert_allclose(e2["valid"][k], e1["valid"][k], rtol=1e-12)
    # host and device-resident loaders agree; the reloaded weights (position bias scaled) changed the position score on BOTH lanes
    np.testing.assert_allclose(two[0][0]["valid"]["esa"], two[1][0]["valid"]["esa"], rtol=1e-12)
    assert abs(two[-1][0]["valid"]["pos"][0] - two[0][0]["valid"]["pos"][0]) > 1e-6


def test_packed_upload_is_bit_identical(sd):
    """spef_eval_submit_host with the packed upload (host threads round the float pixels to BF16 -- the stem's first step -- and half
    the bytes cross the bus; csrc/host_pack.cpp) against the plain float copy: per-image errors bit-identical, sums equal; full and
    ragged batches, more submits than staging slots, a NaN pixel still reaches the decode guard; the FP32 engine never packs."""
    hist = O.ori_histogram(12)[0]
    imgs = [synthetic.synthetic_images(b, seed=40 + i).pin_memory() for i, b in enumerate((8, 5, 8, 1, 7))]
    tgs = [synthetic.synthetic_targets(im.shape[0], seed=i) for i, im in enumerate(imgs)]
    res = {}
    for on in (False, True):
        eng = _engine(sd, "bf16")
        eng.set_ori_histogram(hist)
        eng.set_host_pack(on)
        active, threads = eng.host_pack_info()
        assert active == on and (threads >= 1) == on
        eng.eval_reset()
        outs = []
        for im, tg in zip(imgs, tgs):
            per = torch.empty((im.shape[0], 2), dtype=torch.float32).pin_memory()
            eng.eval_submit_host(im, torch.as_tensor(tg["ori"]).float(), torch.as_tensor(tg["pos"]).float(), per)
            outs.append(per)
        eng.eval_wait()
        res[on] = ([o.numpy().copy() for o in outs], eng.eval_read())
        if on:
            bad = imgs[1].clone().pin_memory()
            bad[2, 1, 100, 7] = float("nan")
            eng.eval_reset()
            eng.eval_submit_host(bad, torch.as_tensor(tgs[1]["ori"]).float(), torch.as_tensor(tgs[1]["pos"]).float(), None)
            eng.eval_wait()
            assert eng.eval_read()[6] >= 1          # the NaN image is counted by the orientation decode guard (sums[6]), as with the plain copy
        eng.close()
    for a, b in zip(res[False][0], res[True][0]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(res[False][1], res[True][1])
    assert res[True][1][3] == sum(im.shape[0] for im in imgs)
    eng = _engine(sd, "fp32")
    assert eng.host_pack_info() == (False, 0)
    eng.close()


def test_conv_pool_kernel_is_bit_identical_to_two_launches(sd, monkeypatch):
    """The last 1x1 conv (320 -> 1280 @8x12) and the global average pool as ONE kernel (csrc/conv_pool.cuh: transposed GEMM, a thread
    owns a channel, the pool is a sum over its registers in the pool kernel's order) against the GEMM + global_mean_kernel launches
    (SPEF_POOL_FUSE=0; each of them is teacher-forced against the oracle above): logits and positions bit-identical for odd, single
    and full batches (the odd image of a pair, CTAs without work, every CTA looping over several image pairs)."""
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SPEF_POOL_FUSE", mode)
        eng = _engine(sd, "bf16", max_batch=256)
        l0 = eng.launch_count()
        res = []
        for b, seed in ((5, 11), (1, 12), (256, 13)):
            x = synthetic.synthetic_images(min(b, 16), seed=seed)
            x = x.repeat((b + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:b].contiguous()
            res.append([t.cpu().numpy() for t in eng.forward(x)])
        outs[mode] = (res, eng.launch_count() - l0)
        eng.close()
    for (o1, p1), (o0, p0) in zip(outs["1"][0], outs["0"][0]):
        np.testing.assert_array_equal(o1, o0)
        np.testing.assert_array_equal(p1, p0)
    assert outs["0"][1] - outs["1"][1] == 3, outs      # one launch fewer per forward


def test_programmatic_dependent_launch_changes_no_bit(sd, monkeypatch):
    """The forward chain is launched with programmatic stream serialization (every tcgen05 / TMA kernel does its set-up, then
    griddepcontrol.wait before it touches the previous kernel's output; csrc/common.cuh) -- with the early trigger for small batches
    (B = 5) and without it for batches that fill the GPU (B = 128 > SPEF_PDL_MAX_BATCH): logits bit-identical to plain stream order
    (SPEF_PDL=0), directly and through the replayed evaluation graph, repeated to give a missing wait a chance to show."""
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SPEF_PDL", mode)
        eng = _engine(sd, "bf16", max_batch=128)
        eng.set_ori_histogram(O.ori_histogram(12)[0])
        res[mode] = []
        for B in (5, 128):
            x = synthetic.synthetic_images(min(B, 8), seed=21).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous()
            tg = synthetic.synthetic_targets(B)
            qt, tt = torch.as_tensor(tg["ori"]).float().cuda(), torch.as_tensor(tg["pos"]).float().cuda()
            outs = [[t.cpu().numpy() for t in eng.forward(x)] for _ in range(4)]
            xd = x.cuda()
            pers = [eng.eval_batch(xd, qt, tt, want_per_image=True).cpu().numpy() for _ in range(5)]   # direct, then captured + replayed
            res[mode].append((outs, pers))
        eng.close()
    for (o1, p1), (o0, p0) in zip(res["1"], res["0"]):
        for a, b in zip(o1, o0):
            np.testing.assert_array_equal(a[0], b[0])
            np.testing.assert_array_equal(a[1], b[1])
        for a, b in zip(p1, p0):
            np.testing.assert_array_equal(a, b)
