"""CPU: pins oracle/indices.py against the golden vectors HF's get_window_index / rot_pos_emb
produced (tests/golden/make_golden.py), the SURVEY §8c known answers, and the live HF module."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, hf_available
from oracle import indices


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "indices_*.npz"))), ids=os.path.basename)
def test_golden(path):
    g = np.load(path)
    win, cu, raw = indices.window_index(g["grid"])
    assert np.array_equal(win, g["window_index"])
    assert np.array_equal(cu, g["cu_dedup"]) and cu.dtype == np.int32
    assert raw == g["cu_raw"].tolist()
    pos = indices.pos_hw(g["grid"])
    inv = 1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))
    freqs = torch.outer(torch.arange(int(pos.max()) + 1, dtype=torch.float), inv)[torch.from_numpy(pos).long()]
    assert np.array_equal(freqs.flatten(1).numpy(), g["rot_pos_emb"])


def test_known_answers_survey_8c():
    win, cu, raw = indices.window_index([[4, 12, 20]])
    assert win[:40].tolist() == [0, 1, 2, 3, 10, 11, 12, 13, 20, 21, 22, 23, 30, 31, 32, 33, 4, 5, 6, 7, 14, 15, 16,
                                 17, 24, 25, 26, 27, 34, 35, 36, 37, 8, 9, 18, 19, 28, 29, 38, 39]
    assert cu[:8].tolist() == [0, 64, 128, 160, 192, 224, 240, 304]
    win, cu, raw = indices.window_index([[8, 32, 32]])
    assert win[:20].tolist() == [0, 1, 2, 3, 16, 17, 18, 19, 32, 33, 34, 35, 48, 49, 50, 51, 4, 5, 6, 7]
    assert len(raw) == 201 and len(cu) == 129
    assert indices.cu_seqlens([[8, 32, 32]]).tolist() == [1024 * i for i in range(9)]
    assert indices.cu_seqlens([[2, 4, 6], [1, 8, 8]]).tolist() == [0, 24, 48, 112]


def test_reverse_is_inverse_permutation():
    win, _, _ = indices.window_index([[3, 26, 46], [1, 4, 6]])
    rev = indices.reverse_index(win)
    assert np.array_equal(win[rev], np.arange(len(win)))
    assert sorted(win.tolist()) == list(range(len(win)))


@pytest.mark.skipif(not hf_available(), reason="transformers not importable")
def test_against_live_hf_module():
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel as V

    cfg = Qwen2_5_VLVisionConfig(depth=1, hidden_size=1280, intermediate_size=3420, num_heads=16,
                                 out_hidden_size=3584, patch_size=14, spatial_merge_size=2, temporal_patch_size=2,
                                 window_size=112, fullatt_block_indexes=[0], hidden_act="silu")
    with torch.device("meta"):
        m = V._from_config(cfg)
    rng = np.random.default_rng(0)
    for _ in range(60):
        grid = [[int(rng.integers(1, 5)), int(rng.integers(1, 24)) * 2, int(rng.integers(1, 24)) * 2]
                for _ in range(int(rng.integers(1, 4)))]
        wi, cu = V.get_window_index(m, torch.tensor(grid))
        win, cud, raw = indices.window_index(grid)
        assert np.array_equal(wi.numpy(), win) and raw == list(cu)
        assert np.array_equal(torch.unique_consecutive(torch.tensor(cu, dtype=torch.int32)).numpy(), cud)
