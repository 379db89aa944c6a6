"""CPU: the resize oracle against the goldens made by the reference's own torchvision call, against torch live, and the
host-side size / frame-count helpers (oracle and product) against tables produced by the reference module itself."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import resize as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "resize_*x*.npz")))


def test_goldens_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_oracle_matches_reference_call(path):
    g = np.load(path)
    oh, ow = g["out"].shape[-2:]
    raw = R.resize_bicubic_aa_f32(g["frames"], oh, ow)
    assert raw.dtype == np.float32 and np.array_equal(raw.view(np.uint32), g["raw"].view(np.uint32))   # float bits
    assert np.array_equal(R.resize_frames(g["frames"], oh, ow), g["out"])


@pytest.mark.parametrize("shape,size", [((1, 1, 33, 47), (28, 28)), ((1, 2, 28, 28), (56, 84)), ((2, 1, 120, 64), (28, 64)),
                                        ((1, 1, 31, 400), (31, 28))])
def test_oracle_matches_torch_live(shape, size):
    x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(sum(shape)))
    ref = F.interpolate(x.float(), size=size, mode="bicubic", antialias=True, align_corners=False).numpy()
    assert np.array_equal(R.resize_bicubic_aa_f32(x.numpy(), *size), ref)


def test_weights_known_answers():
    # identity: every output index has weight 1 on its own input (cubic(0) = 1, cubic(+-1) = 0)
    for lo, w in R.axis_weights(12, 12):
        assert np.isclose(w.sum(), 1) and w.max() == 1 and np.count_nonzero(w) == 1
    # 2x reduction: 8 taps in the interior, symmetric, summing to 1
    lo, w = R.axis_weights(64, 32)[10]
    assert len(w) == 8 and lo == 17 and np.allclose(w, w[::-1]) and np.isclose(w.sum(), 1, atol=1e-6)
    # by hand: cubic(.25, .75, 1.25, 1.75) = .8671875, .2265625, -.0703125, -.0234375; their sum over both sides is 2
    assert np.array_equal(w[:4], np.float32([-0.01171875, -0.03515625, 0.11328125, 0.43359375]))


def _helpers():
    return np.load(os.path.join(GOLD, "resize_helpers.npz"))


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_smart_resize_table(impl):
    if impl == "product":
        from video_r2_b200 import preprocess as P
        fn = P.smart_resize
    else:
        fn = R.smart_resize
    rows = _helpers()["sizes"]
    assert len(rows) > 200
    for h, w, lo, hi, eh, ew in rows.tolist():
        assert fn(h, w, 28, lo, hi) == (eh, ew), (h, w, lo, hi)
    with pytest.raises(ValueError):
        fn(10, 2001)


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_smart_nframes_table(impl):
    if impl == "product":
        from video_r2_b200 import preprocess as P
        fn = P.smart_nframes
    else:
        fn = R.smart_nframes
    for total, vfps, fps, arg, expect in _helpers()["frames"].tolist():
        total, arg, expect = int(total), int(arg), int(expect)
        kw = {"nframes": arg} if fps < 0 else {"fps": fps, "max_frames": arg}
        if expect < 0:
            with pytest.raises(ValueError):
                fn(total, vfps, **kw)
        else:
            assert fn(total, vfps, **kw) == expect, (total, vfps, fps, arg)


def test_video_target_size():
    from video_r2_b200 import preprocess as P
    # 720p, 32 frames: capped by VIDEO_MAX_PIXELS = 768*28*28; eval's max_pixels=151200 gives the 16:9 grid of SURVEY §8
    assert P.video_target_size(32, 720, 1280) == R.smart_resize(720, 1280, 28, R.VIDEO_MIN_PIXELS, 768 * 28 * 28)
    assert P.video_target_size(128, 720, 1280, max_pixels=151200) == (280, 504)
    assert int(_helpers()["total_pixels"]) == P.VIDEO_TOTAL_PIXELS
