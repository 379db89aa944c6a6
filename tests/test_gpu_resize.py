"""GPU: vr2_resize_frames through the C ABI against the goldens made by the reference's own call (torchvision tensor
resize, bicubic, antialias), against the oracle on seeded inputs, and — at real clip sizes — against torch's CPU
kernel run on the box (torch ships in the image; /root/reference is not needed).  Bytes must be identical."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import resize as R

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "resize_*x*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[7:-4] for p in CASES])
def test_goldens(path):
    import video_r2_b200 as vr2
    g = np.load(path)
    frames = torch.from_numpy(g["frames"]).cuda()
    if bool(g["hwc"]):                                     # decoder layout: [T,H,W,C] memory viewed as [T,C,H,W]
        frames = frames.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        assert not frames.is_contiguous()
    oh, ow = g["out"].shape[-2:]
    out = vr2.resize_frames(frames, (oh, ow))
    assert out.dtype == torch.uint8 and out.is_contiguous()
    assert np.array_equal(out.cpu().numpy(), g["out"])
    outf = vr2.resize_frames(frames, (oh, ow), out_dtype=torch.float32)
    assert np.array_equal(outf.cpu().numpy(), g["out"].astype(np.float32))      # `.float()` of the reference


@pytest.mark.parametrize("shape,size", [((1, 1, 33, 47), (28, 28)), ((2, 3, 28, 28), (56, 84)), ((1, 1, 31, 400), (31, 28)),
                                        ((1, 2, 600, 40), (28, 56)), ((3, 3, 97, 131), (97, 131)), ((1, 1, 4, 4), (28, 28)),
                                        ((1, 1, 2000, 30), (14, 30))])
def test_against_oracle(shape, size):
    import video_r2_b200 as vr2
    x = torch.randint(0, 256, shape, dtype=torch.uint8, generator=torch.Generator().manual_seed(sum(shape)))
    want = R.resize_frames(x.numpy(), *size)
    got = vr2.resize_frames(x.cuda(), size)
    assert np.array_equal(got.cpu().numpy(), want)


def test_real_clip_sizes_against_torch_cpu():
    """720p and 1080p sources to the sizes fetch_video picks (video_target_size), 4 frames each, plus an upscale."""
    import video_r2_b200 as vr2
    for (H, W, T, max_pixels) in ((720, 1280, 4, None), (1080, 1920, 4, 151200), (240, 320, 4, None)):
        g = torch.Generator().manual_seed(H)
        x = torch.randint(0, 256, (T, H, W, 3), dtype=torch.uint8, generator=g).permute(0, 3, 1, 2)
        # smooth half of each frame so that it is not all noise
        x[:, :, :, : W // 2] = (torch.arange(W // 2)[None, None, None, :] * 255 // (W // 2)).to(torch.uint8)
        size = vr2.video_target_size(32, H, W, max_pixels=max_pixels)
        ref = F.interpolate(x.float(), size=size, mode="bicubic", antialias=True, align_corners=False)
        ref = ref.clamp(0, 255).round().to(torch.uint8)
        got = vr2.resize_frames(x.cuda(), size)
        assert torch.equal(got.cpu(), ref), (H, W, size, (got.cpu() != ref).sum().item())


def test_preprocess_video_matches_resize_then_processor():
    """frames -> pixel_values through the two kernels == oracle resize followed by the patchify oracle."""
    import video_r2_b200 as vr2
    from oracle import patchify as OP
    x = torch.randint(0, 256, (4, 3, 120, 200), dtype=torch.uint8, generator=torch.Generator().manual_seed(9))
    size = vr2.smart_resize(120, 200, 28, 4 * 28 * 28, 16 * 28 * 28)
    pix, grid = vr2.preprocess_video(x.cuda(), size=size, out_dtype=torch.float32)
    from video_r2_b200.preprocess import OPENAI_CLIP_MEAN, OPENAI_CLIP_STD
    want, wgrid = OP.patchify(R.resize_frames(x.numpy(), *size), OPENAI_CLIP_MEAN, OPENAI_CLIP_STD)
    assert list(grid) == list(wgrid)
    assert np.array_equal(pix.cpu().numpy(), np.asarray(want, dtype=np.float32))


def test_argument_errors():
    import video_r2_b200 as vr2
    with pytest.raises(ValueError):
        vr2.resize_frames(torch.zeros(1, 3, 8, 8, dtype=torch.uint8), (4, 4))              # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        vr2.resize_frames(torch.zeros(1, 3, 8, 8, device="cuda"), (4, 4))                   # not uint8
    with pytest.raises(ValueError):
        vr2.resize_frames(torch.zeros(1, 3, 8, 8, dtype=torch.uint8, device="cuda"), (0, 4))                  # VR2_ERR_ARG
