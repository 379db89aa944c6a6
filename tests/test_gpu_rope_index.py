"""GPU: vr2.get_rope_index is bit-exact against the oracle restatement of transformers 4.51.1."""
import numpy as np
import pytest
import torch

from oracle import rope_index as R
from rope_cases import VID, VS, build_row, case_suite, pad_batch

pytestmark = pytest.mark.gpu


def _gpu(kw):
    out = {}
    for k, v in kw.items():
        out[k] = v.cuda() if isinstance(v, torch.Tensor) else v
    return out


def _run_both(kw, int_cast):
    import video_r2_b200 as vr2
    cfg = vr2.RopeConfig(spg_int_cast=int_cast)
    g = _gpu(kw)
    pos, delta = vr2.get_rope_index(g["input_ids"], g.get("image_grid_thw"), g.get("video_grid_thw"),
                                    g.get("second_per_grid_ts"), g.get("attention_mask"), config=cfg)
    rpos, rdelta = R.get_rope_index(kw["input_ids"], kw.get("image_grid_thw"), kw.get("video_grid_thw"),
                                    kw.get("second_per_grid_ts"), kw.get("attention_mask"), spg_int_cast=int_cast)
    return pos.cpu(), delta.cpu(), rpos, rdelta


@pytest.mark.parametrize("int_cast", [True, False], ids=["int64cast", "float32mul"])
@pytest.mark.parametrize("name,kw", case_suite(), ids=[c[0] for c in case_suite()])
def test_bit_exact_vs_oracle(name, kw, int_cast):
    pos, delta, rpos, rdelta = _run_both(kw, int_cast)
    assert pos.dtype == torch.int64 and pos.shape == rpos.shape and delta.shape == rdelta.shape
    assert torch.equal(pos, rpos), name
    assert torch.equal(delta, rdelta.to(torch.int64)), name


def test_fp32_truncation_known_answers():
    import video_r2_b200 as vr2
    ids = torch.tensor([[VS] + [VID] * 12 + [151653]]).cuda()
    grid = torch.tensor([[12, 2, 2]]).cuda()
    for spg, ef, ei in [(0.7, [0, 1, 2, 4, 5, 7, 8, 9, 11, 12, 14, 15], [0] * 12),
                        (9.375, [0, 18, 37, 56, 75, 93, 112, 131, 150, 168, 187, 206], [18 * k for k in range(12)])]:
        pos, _ = vr2.get_rope_index(ids, None, grid, torch.tensor([spg]).cuda(), None, config=vr2.RopeConfig(spg_int_cast=False))
        assert (pos[0, 0, 1:13] - 1).tolist() == ef
        pos, _ = vr2.get_rope_index(ids, None, grid, [spg], None, config=vr2.RopeConfig(spg_int_cast=True))
        assert (pos[0, 0, 1:13] - 1).tolist() == ei


@pytest.mark.parametrize("int_cast", [True, False])
def test_config2_sequence_and_long_batch(int_cast):
    """BASELINE config 2: a 32-frame 448x448 clip = 4096 video tokens + text; plus a ragged 6-row batch."""
    rng = np.random.default_rng(5)
    g = [16, 32, 32]
    row = build_row([(VID, g)], rng, lead=32, gap=0, tail=64)
    kw = dict(input_ids=torch.tensor([row]), video_grid_thw=torch.tensor([g]), second_per_grid_ts=torch.tensor([2 / 0.7]))
    pos, delta, rpos, rdelta = _run_both(kw, int_cast)
    assert torch.equal(pos, rpos) and torch.equal(delta, rdelta)
    grids = [[int(rng.integers(1, 9)), int(rng.integers(2, 12)) * 2, int(rng.integers(2, 12)) * 2] for _ in range(6)]
    rows = [build_row([(VID, gg)], rng, lead=int(rng.integers(0, 40)), gap=int(rng.integers(0, 5)), tail=int(rng.integers(0, 50))) for gg in grids]
    for side in ("left", "right"):
        ids, mask = pad_batch(rows, side)
        kw = dict(input_ids=ids, attention_mask=mask, video_grid_thw=torch.tensor(grids),
                  second_per_grid_ts=torch.tensor([2 / f for f in (2.0, 1.0, 0.2133, 0.7, 3.0, 1.7)]))
        pos, delta, rpos, rdelta = _run_both(kw, int_cast)
        assert torch.equal(pos, rpos) and torch.equal(delta, rdelta)


@pytest.mark.parametrize("int_cast", [True, False], ids=["int64cast", "float32mul"])
@pytest.mark.parametrize("seed", range(12))
def test_random_batches_bit_exact(seed, int_cast):
    """Seeded fuzz: batches of 1-6 rows, each with 0-4 vision items (images and videos mixed in any order, text-only
    rows included), random even grids, random text gaps (also zero: items back to back), left or right padding or no
    mask, fractional seconds-per-grid as tensor / list / None — position ids and deltas equal the oracle's, bit for bit."""
    from rope_cases import IMG
    rng = np.random.default_rng(1000 + seed)
    B = int(rng.integers(1, 7))
    rows, img_grids, vid_grids = [], [], []
    for _ in range(B):
        items = []
        for _ in range(int(rng.integers(0, 5))):
            kind = IMG if rng.random() < 0.4 else VID
            g = [1 if kind == IMG else int(rng.integers(1, 7)), int(rng.integers(1, 9)) * 2, int(rng.integers(1, 9)) * 2]
            items.append((kind, g))
            (img_grids if kind == IMG else vid_grids).append(g)
        rows.append(build_row(items, rng, lead=int(rng.integers(0, 12)), gap=int(rng.integers(0, 4)), tail=int(rng.integers(0, 9))))
        if not rows[-1]:
            rows[-1] = [1234]
    mode = ["left", "right", "none"][seed % 3]
    if mode == "none":   # no attention mask: rows must have one length
        S = max(len(r) for r in rows)
        rows = [r + rng.integers(1000, 5000, size=S - len(r)).tolist() for r in rows]
        kw = dict(input_ids=torch.tensor(rows))
    else:
        ids, mask = pad_batch(rows, mode)
        kw = dict(input_ids=ids, attention_mask=mask)
    if img_grids:
        kw["image_grid_thw"] = torch.tensor(img_grids)
    if vid_grids:
        kw["video_grid_thw"] = torch.tensor(vid_grids)
        spg = [float(2.0 / rng.choice([2.0, 1.0, 0.7, 0.2133, 3.0, 0.5, 1.7])) for _ in vid_grids]
        how = seed % 4
        if how == 0:
            kw["second_per_grid_ts"] = torch.tensor(spg, dtype=torch.float32)
        elif how == 1:
            kw["second_per_grid_ts"] = spg
        elif how == 2:
            kw["second_per_grid_ts"] = torch.tensor([float(round(v)) for v in spg], dtype=torch.float32)
    if not img_grids and not vid_grids and mode == "none":
        kw["attention_mask"] = torch.ones_like(kw["input_ids"])
    pos, delta, rpos, rdelta = _run_both(kw, int_cast)
    assert torch.equal(pos, rpos), (seed, mode)
    assert torch.equal(delta, rdelta.to(torch.int64)), (seed, mode)


# ---- the CUDA kernel against the fixtures of the two live implementations directly (no oracle in between)
import json
import os

_G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(_G, "rope_vllm.json")) as _f:
    _VLLM = [r for r in json.load(_f)["rows"]
             if r["name"].startswith(("exact_spg", "config2", "docstring", "video_nomask", "image_only", "video_spg_none",
                                      "video_list_spg"))]
with open(os.path.join(_G, "rope_hf55.json")) as _f:
    _HF55 = json.load(_f)["batches"]


@pytest.mark.parametrize("row", _VLLM, ids=[r["name"] for r in _VLLM])
def test_kernel_vs_installed_vllm_fixture(row):
    """Rows whose seconds-per-grid values are exact in binary: fp32 == fp64, so vLLM's output is the expected output of
    the multiply-then-truncate mode, element for element."""
    import video_r2_b200 as vr2
    ids = torch.tensor([row["tokens"]]).cuda()
    img = torch.tensor(row["image_grid_thw"]).cuda() if row["image_grid_thw"] else None
    vid = torch.tensor(row["video_grid_thw"]).cuda() if row["video_grid_thw"] else None
    spg = None if row["second_per_grid_ts"] is None else torch.tensor(row["second_per_grid_ts"], dtype=torch.float32).cuda()
    cfg = vr2.RopeConfig(spg_int_cast=False, tokens_per_second=row["tokens_per_second"])
    pos, delta = vr2.get_rope_index(ids, img, vid, spg, None, config=cfg)
    ref = torch.tensor(row["pos"])
    assert torch.equal(pos[:, 0].cpu(), ref), row["name"]
    assert int(delta.cpu()) == row["delta"]


@pytest.mark.parametrize("b", _HF55, ids=[b["name"] for b in _HF55])
def test_kernel_vs_installed_transformers_fixture(b):
    import video_r2_b200 as vr2
    ids, mask = torch.tensor(b["input_ids"]), torch.tensor(b["attention_mask"])
    img = torch.tensor(b["image_grid_thw"]).cuda() if b["image_grid_thw"] else None
    vid = torch.tensor(b["video_grid_thw"]).cuda() if b["video_grid_thw"] else None
    for int_cast in (True, False):
        cfg = vr2.RopeConfig(spg_int_cast=int_cast, tokens_per_second=b["tokens_per_second"])
        pos, delta = vr2.get_rope_index(ids.cuda(), img, vid, None, mask.cuda(), config=cfg)
        pos, delta = pos.cpu(), delta.cpu()
        ref = torch.tensor(b["pos"])
        m = mask.bool()
        assert torch.equal(pos[:, m], ref[:, m])
        assert (pos[:, ~m] == 1).all()
        assert torch.equal(delta.flatten() + (mask.shape[1] - mask.sum(1)), torch.tensor(b["delta"]))
