"""GPU: vr2_patchify against the oracle (and so against the HF video processor): bit-exact in fp32, and equal to the
processor's output cast to bf16 (`.type(self.visual.dtype)`, monkey_patch_forward.py:391)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import patchify as P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "patchify_*.npz"))), ids=os.path.basename)
def test_patchify_matches_hf_golden(path):
    import video_r2_b200 as vr2
    g = np.load(path)
    frames = torch.from_numpy(g["frames"]).cuda()
    out32, grid = vr2.patchify(frames, g["mean"].tolist(), g["std"].tolist(), float(g["rescale"]), out_dtype=torch.float32)
    assert grid == g["grid"].tolist()
    assert torch.equal(out32.cpu(), torch.from_numpy(g["out"]))
    out16, _ = vr2.patchify(frames, g["mean"].tolist(), g["std"].tolist(), float(g["rescale"]))
    assert out16.dtype == torch.bfloat16 and torch.equal(out16.cpu(), torch.from_numpy(g["out"]).to(torch.bfloat16))
    # float frames (Video-R2 resizes and calls .float() before the processor): same numbers
    outf, _ = vr2.patchify(frames.float(), g["mean"].tolist(), g["std"].tolist(), float(g["rescale"]), out_dtype=torch.float32)
    assert torch.equal(outf, out32)


def test_patchify_full_size_clip_and_errors():
    import video_r2_b200 as vr2
    frames = torch.randint(0, 256, (32, 3, 448, 448), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    ref, grid = P.patchify(frames.numpy(), vr2.preprocess.OPENAI_CLIP_MEAN, vr2.preprocess.OPENAI_CLIP_STD)
    out, g = vr2.patchify(frames.cuda(), out_dtype=torch.float32)
    assert g == grid == [16, 32, 32] and torch.equal(out.cpu(), torch.from_numpy(ref))
    with pytest.raises(ValueError):
        vr2.patchify(torch.zeros(2, 3, 30, 28, dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        vr2.patchify(torch.zeros(2, 3, 28, 28, dtype=torch.int32, device="cuda"))
