"""CPU: pins oracle/rope_index.py (transformers-4.51.1 get_rope_index restatement) with the
known answers listed in SURVEY §8c and an independent second implementation."""
import numpy as np
import pytest
import torch

from oracle import rope_index as R
from rope_cases import IMG, VE, VID, VS, case_suite


def test_hf_docstring_example():
    # HF docstring: 3x2x2 merged grid, interval 50 (tokens_per_second 25, 2 s per grid) ->
    # t = 0,0,0,0,50,...,100; h = 0,0,1,1; w = 0,1,0,1; following text starts at 101.
    # With the vision_start token in front (text_len 1) everything shifts by 1.
    ids = torch.tensor([[VS] + [VID] * 12 + [11, 12, 13, 14, 15]])
    for int_cast in (True, False):
        pos, delta = R.get_rope_index(ids, None, torch.tensor([[3, 4, 4]]), torch.tensor([2.0]), None,
                                      tokens_per_second=25, spg_int_cast=int_cast)
        assert pos[0, 0].tolist() == [0] + [1] * 4 + [51] * 4 + [101] * 4 + [102, 103, 104, 105, 106]
        assert pos[1, 0, 1:13].tolist() == [1, 1, 2, 2] * 3
        assert pos[2, 0, 1:13].tolist() == [1, 2, 1, 2] * 3
        assert delta.tolist() == [[107 - 18]]


def test_hand_derived_survey_kat():
    ids = torch.tensor([[1, 2, VS] + [VID] * 12 + [VE, 5, 6]])
    pos, delta = R.get_rope_index(ids, None, torch.tensor([[3, 4, 4]]), torch.tensor([1.0]), None)
    assert pos[0, 0].tolist() == [0, 1, 2, 3, 3, 3, 3, 5, 5, 5, 5, 7, 7, 7, 7, 8, 9, 10]
    assert pos[1, 0, 3:15].tolist() == [3, 3, 4, 4] * 3
    assert pos[2, 0, 3:15].tolist() == [3, 4, 3, 4] * 3
    assert delta.tolist() == [[-7]]


@pytest.mark.parametrize("spg,expect_float,expect_int", [
    (0.7, [0, 1, 2, 4, 5, 7, 8, 9, 11, 12, 14, 15], [0] * 12),
    (9.375, [0, 18, 37, 56, 75, 93, 112, 131, 150, 168, 187, 206], [18 * k for k in range(12)]),
])
def test_temporal_truncation_modes(spg, expect_float, expect_int):
    ids = torch.tensor([[VS] + [VID] * 12 + [VE]])
    grid = torch.tensor([[12, 2, 2]])
    for as_list in (False, True):
        s = [spg] if as_list else torch.tensor([spg])
        pos, _ = R.get_rope_index(ids, None, grid, s, None, spg_int_cast=False)
        assert (pos[0, 0, 1:13] - 1).tolist() == expect_float  # float32 multiply then .long()
        pos, _ = R.get_rope_index(ids, None, grid, s, None, spg_int_cast=True)
        assert (pos[0, 0, 1:13] - 1).tolist() == expect_int    # int64 cast first


def test_padding_positions_are_one_and_delta_uses_padded_length():
    cases = dict(case_suite())
    kw = cases["mixed_batch_leftpad"]
    pos, delta = R.get_rope_index(kw["input_ids"], kw["image_grid_thw"], kw["video_grid_thw"], kw["second_per_grid_ts"],
                                  kw["attention_mask"])
    m = kw["attention_mask"]
    assert (pos[:, m == 0] == 1).all()
    S = m.shape[1]
    for b in range(m.shape[0]):
        assert delta[b, 0].item() == pos[:, b, m[b] == 1].max().item() + 1 - S


@pytest.mark.parametrize("name,kw", case_suite(), ids=[c[0] for c in case_suite()])
def test_against_independent_second_implementation(name, kw):
    if "image_grid_thw" not in kw and "video_grid_thw" not in kw:
        pytest.skip("text-only branch has no vLLM counterpart")
    spg = kw.get("second_per_grid_ts")
    ids, mask = kw["input_ids"], kw.get("attention_mask")
    # float mode vs float64 second opinion: only where fp32 and fp64 truncation agree (checked per video)
    pos, delta = R.get_rope_index(ids, kw.get("image_grid_thw"), kw.get("video_grid_thw"), spg, mask, spg_int_cast=False)
    ii = vi = 0
    for b in range(ids.shape[0]):
        row = ids[b][mask[b].bool()] if mask is not None else ids[b]
        toks = row.tolist()
        n_i = sum(1 for j, v in enumerate(toks[:-1]) if v == VS and toks[j + 1] == IMG)
        n_v = sum(1 for j, v in enumerate(toks[:-1]) if v == VS and toks[j + 1] == VID)
        img = kw["image_grid_thw"][ii:ii + n_i].tolist() if n_i else []
        vid = kw["video_grid_thw"][vi:vi + n_v].tolist() if n_v else []
        s = None
        if spg is not None and n_v:
            s = [float(torch.as_tensor(spg, dtype=torch.float32)[vi + k]) for k in range(n_v)]
        ii, vi = ii + n_i, vi + n_v
        ref, ref_delta = R.vllm_style(toks, img, vid, s)
        got = pos[:, b, mask[b].bool()] if mask is not None else pos[:, b]
        # fp32-vs-fp64 disagreement only ever moves a temporal index by 1; h/w rows and text must match exactly
        assert np.array_equal(got[1:].numpy(), ref[1:])
        assert np.abs(got[0].numpy() - ref[0]).max() <= 1
