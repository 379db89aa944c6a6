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


# ---------------------------------------------------------------------------------------------------------------
# Pins from the two LIVE implementations in the image (fixtures written by tests/golden/make_golden.py rope):
# the installed vLLM's get_mrope_input_positions and the installed transformers-5.5.0 get_rope_index.
# ---------------------------------------------------------------------------------------------------------------
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def _agreeing_temporal_mask(row):
    """Per token: True where the fp32 product chain of 4.51.1 (torch: int64 * float32 -> float32) and vLLM's float64
    product truncate to the same integer — computed here with numpy, independently of the oracle."""
    toks = np.asarray(row["tokens"])
    ok = np.ones(len(toks), dtype=bool)
    spg = row["second_per_grid_ts"]
    if not row["video_grid_thw"]:
        return ok, 0
    tps = row["tokens_per_second"]
    pads = np.flatnonzero(toks == VID)
    at = 0
    disagree = 0
    for vi, (t, h, w) in enumerate(row["video_grid_thw"]):
        per = (h // 2) * (w // 2)
        s = 1.0 if spg is None else spg[vi]
        k = np.arange(t)
        f32 = np.trunc((k.astype(np.float32) * np.float32(s)).astype(np.float32) * np.float32(tps)).astype(np.int64)
        f64 = (k * float(s) * tps).astype(np.int64)
        same = np.repeat(f32 == f64, per)
        ok[pads[at:at + t * per]] = same
        disagree += int((~same).sum())
        at += t * per
    # a disagreement in a video's LAST slice can shift everything behind it (text resumes at max + 1)
    if disagree:
        first_bad = int(np.flatnonzero(~ok)[0])
        ok[first_bad:] = False
    return ok, disagree


VLLM_ROWS = _load("rope_vllm.json")["rows"]


@pytest.mark.parametrize("row", VLLM_ROWS, ids=[r["name"] for r in VLLM_ROWS])
def test_pinned_by_installed_vllm(row):
    """multiply-then-truncate mode == vLLM 0.22's real get_mrope_input_positions, element for element, wherever the fp32
    and fp64 products truncate alike (every element for exactly representable seconds-per-grid values)."""
    ids = torch.tensor([row["tokens"]])
    img = torch.tensor(row["image_grid_thw"]) if row["image_grid_thw"] else None
    vid = torch.tensor(row["video_grid_thw"]) if row["video_grid_thw"] else None
    spg = None if row["second_per_grid_ts"] is None else torch.tensor(row["second_per_grid_ts"], dtype=torch.float32)
    pos, delta = R.get_rope_index(ids, img, vid, spg, None, tokens_per_second=row["tokens_per_second"], spg_int_cast=False)
    ref = np.asarray(row["pos"])
    ok, disagree = _agreeing_temporal_mask(row)
    got = pos[:, 0].numpy()
    assert np.array_equal(got[:, ok], ref[:, ok]), row["name"]
    assert np.abs(got - ref).max() <= 1          # a truncation disagreement moves an index by at most one
    if disagree == 0:
        assert int(delta) == row["delta"]
    if row["name"].startswith(("exact_spg", "config2", "docstring", "video_nomask", "image_only", "video_spg_none")):
        assert disagree == 0 and ok.all()        # these rows are pinned completely


def test_vllm_fixture_covers_every_branch():
    names = " ".join(r["name"] for r in VLLM_ROWS)
    for need in ("mixed_batch_leftpad", "grpo_repeat", "image_only", "video_frac_spg", "exact_spg_9.375", "config2"):
        assert need in names
    assert sum(len(r["tokens"]) for r in VLLM_ROWS) > 5000


HF55 = _load("rope_hf55.json")["batches"]


@pytest.mark.parametrize("b", HF55, ids=[b["name"] for b in HF55])
def test_pinned_by_installed_transformers_on_t1_items(b):
    """transformers 5.5.0's live get_rope_index on rows whose vision items all have t == 1 (where its arithmetic still
    coincides with 4.51.1): unmasked positions equal; its delta uses the unpadded length, 4.51.1 the padded one."""
    ids, mask = torch.tensor(b["input_ids"]), torch.tensor(b["attention_mask"])
    img = torch.tensor(b["image_grid_thw"]) if b["image_grid_thw"] else None
    vid = torch.tensor(b["video_grid_thw"]) if b["video_grid_thw"] else None
    for int_cast in (True, False):
        pos, delta = R.get_rope_index(ids, img, vid, None, mask, tokens_per_second=b["tokens_per_second"], spg_int_cast=int_cast)
        ref = torch.tensor(b["pos"])
        m = mask.bool()
        assert torch.equal(pos[:, m], ref[:, m])
        assert (pos[:, ~m] == 1).all() and (ref[:, ~m] == 0).all()       # the documented padding difference
        pad = mask.shape[1] - mask.sum(1)
        assert torch.equal(delta.flatten() + pad, torch.tensor(b["delta"]))
