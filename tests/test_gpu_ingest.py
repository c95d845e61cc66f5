"""GPU parity of the input side of the path: camera frames -> Resize -> ToTensor through the C ABI
(spef_resize_frames) against goldens produced by the reference's own transform (torchvision + Pillow) and against the
CPU oracle.  8-bit / fixed-point work: bit-exact."""
import json

import numpy as np
import pytest
import torch

from oracle import spef_oracle as O
from spef_b200.tools import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from spef_b200.engine import Engine
    return Engine(240, 384, 1728, 3, False, "bf16", 4)


def test_resize_vs_reference_golden(eng, golden):
    from spef_b200.engine import Engine
    g = golden("resize")
    sq = None
    for name, kw, size in json.loads(str(g["cases"])):
        frames = synthetic.synthetic_frames(**kw)
        e = eng
        if tuple(size) != (240, 384):   # the reference's other default, img_size (240, 240): src/data/datasets/speed.py:38
            sq = sq or Engine(size[0], size[1], 1728, 3, False, "bf16", 1)
            e = sq
        got = e.resize_frames(torch.from_numpy(frames), torch.uint8).cpu().numpy()
        want = g[name]
        if want.shape[1] == 1:
            want = np.repeat(want, 3, axis=1)   # convert("RGB") of a grey frame
        np.testing.assert_array_equal(got, want, err_msg=name)
        f32 = e.resize_frames(torch.from_numpy(frames), torch.float32).cpu().numpy()
        np.testing.assert_array_equal(f32, want.astype(np.float32) / np.float32(255), err_msg=name + " (ToTensor)")
        if name == "speed_1200x1920":
            np.testing.assert_array_equal(f32[0, :, 100, :], g["speed_1200x1920_f32_row100"])


@pytest.mark.parametrize("shape", [(37, 53), (1200, 1920), (2400, 3840), (240, 5000), (3000, 200), (1, 1)])
def test_resize_vs_oracle_ragged_sizes(eng, shape):
    """Sizes the goldens do not hold: tiny / very wide / very tall frames, > 12 horizontal taps (generic tap loop),
    > 8x vertical reduction (smaller row bands), up-scaling in one axis and down-scaling in the other."""
    for ch in (1, 3):
        frames = synthetic.synthetic_frames(2, shape[0], shape[1], ch, seed=shape[0] + ch, kind="noise")
        got = eng.resize_frames(torch.from_numpy(frames), torch.uint8).cpu().numpy()
        np.testing.assert_array_equal(got, O.resize_frames(frames, 240, 384))


def test_resize_properties_full_batch(eng):
    """Size-independent properties at the engine's full batch: a constant frame stays constant (taps sum to 2^22 within
    rounding), slots are independent, the result is deterministic, and a grey frame equals the same frame sent as RGB."""
    frames = synthetic.synthetic_frames(4, 1200, 1920, 1, seed=3, kind="speed")
    a = eng.resize_frames(torch.from_numpy(frames), torch.uint8)
    b = eng.resize_frames(torch.from_numpy(frames), torch.uint8)
    assert torch.equal(a, b)
    one = eng.resize_frames(torch.from_numpy(frames[2:3]), torch.uint8)
    assert torch.equal(one[0], a[2])
    rgb = np.repeat(frames[..., None], 3, axis=-1)
    assert torch.equal(eng.resize_frames(torch.from_numpy(rgb), torch.uint8), a)
    for v in (0, 1, 127, 255):
        c = eng.resize_frames(torch.full((1, 1200, 1920), v, dtype=torch.uint8), torch.uint8)
        assert int(c.min()) == v and int(c.max()) == v


def test_frames_to_pose_matches_tensor_path(eng):
    """Frames -> resize (uint8) -> predict equals predict on the float tensor the reference's loader would have built."""
    sd = synthetic.synthetic_state_dict(1728, 3)
    eng.load_state_dict(sd)
    eng.set_ori_histogram(O.ori_histogram(12)[0])
    frames = synthetic.synthetic_frames(4, 1200, 1920, 1, seed=8, kind="noise")
    x_f32 = torch.from_numpy(O.to_tensor(O.resize_frames(frames, 240, 384)))
    eng.set_image_dtype(torch.float32)
    want = eng.predict(x_f32.cuda())
    eng.set_image_dtype(torch.uint8)
    got = eng.predict(eng.resize_frames(torch.from_numpy(frames)))
    eng.set_image_dtype(torch.float32)
    assert torch.equal(got["ori"], want["ori"]) and torch.equal(got["pos"], want["pos"])


def test_resize_rejects_bad_arguments(eng):
    with pytest.raises(ValueError):
        eng.resize_frames(torch.zeros(1, 10, 10, dtype=torch.float32))
    with pytest.raises(ValueError):
        eng.resize_frames(torch.zeros(1, 10, 10, 2, dtype=torch.uint8))


def test_frame_transform_mirror(eng, golden):
    """spef_b200.data.FrameTransform = the reference's Compose([Resize(img_size), ToTensor()]) for a batch of frames."""
    from spef_b200.data import FrameTransform
    g = golden("resize")
    frames = synthetic.synthetic_frames(batch=2, height=1200, width=1920, channels=1, seed=11, kind="speed")
    tf = FrameTransform(eng, (240, 384))
    x = tf(frames).cpu().numpy()
    assert x.dtype == np.float32 and x.shape == (2, 3, 240, 384)
    np.testing.assert_array_equal(x[0, :, 100, :], g["speed_1200x1920_f32_row100"])
    np.testing.assert_array_equal(tf(frames[0]).cpu().numpy()[0], x[0])           # a single [H,W] frame
    u8 = FrameTransform(eng, (240, 384), torch.uint8)(torch.from_numpy(frames).cuda())
    np.testing.assert_array_equal(u8.cpu().numpy()[:, :1], g["speed_1200x1920"])
    with pytest.raises(ValueError):
        FrameTransform(eng, (240, 240))
