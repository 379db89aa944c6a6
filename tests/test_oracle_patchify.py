"""CPU: the patchify oracle against golden outputs of the HF video processor (tests/golden/patchify_*.npz, written
by tests/golden/make_golden.py) and, when importable, against the live processor on a fresh input."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import patchify as P


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "patchify_*.npz"))), ids=os.path.basename)
def test_oracle_matches_hf_golden(path):
    g = np.load(path)
    out, grid = P.patchify(g["frames"], g["mean"], g["std"], float(g["rescale"]))
    assert grid == g["grid"].tolist()
    assert out.dtype == np.float32 and np.array_equal(out, g["out"])          # bit-exact fp32


def test_oracle_matches_live_processor():
    try:
        from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor
        proc = Qwen2VLVideoProcessor(do_resize=False)
    except Exception as e:  # torchvision / transformers missing
        pytest.skip(f"HF video processor unavailable: {e!r}")
    v = torch.randint(0, 256, (7, 3, 84, 56), dtype=torch.uint8, generator=torch.Generator().manual_seed(11))
    ref = proc(videos=[v], return_tensors="pt", do_resize=False, do_sample_frames=False)
    out, grid = P.patchify(v.numpy(), proc.image_mean, proc.image_std, proc.rescale_factor)
    assert grid == ref["video_grid_thw"][0].tolist()
    assert np.array_equal(out, ref["pixel_values_videos"].numpy())
