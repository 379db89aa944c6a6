"""GPU: the index plan and the row gathers are bit-exact against the oracle and the golden vectors."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import indices as O

pytestmark = pytest.mark.gpu

GRIDS = [
    [[1, 2, 2]], [[1, 28, 28]], [[4, 12, 20]], [[8, 32, 32]], [[3, 26, 46]], [[2, 24, 42]], [[64, 24, 42]],
    [[2, 10, 14], [1, 4, 6], [3, 26, 46], [1, 8, 8]], [[1, 2, 46]], [[5, 46, 2]], [[16, 32, 32]],
]


def _plan(grid):
    import video_r2_b200 as vr2
    return vr2.Plan(vr2.VisionConfig(), grid, "cuda:0")


@pytest.mark.parametrize("grid", GRIDS, ids=[str(g) for g in GRIDS])
def test_plan_bit_exact_vs_oracle(grid):
    ex = _plan(grid).export()
    win, cu_win, _ = O.window_index(grid)
    assert ex["window_index"].dtype == torch.int64 and ex["cu_window_seqlens"].dtype == torch.int32
    assert np.array_equal(ex["window_index"].cpu().numpy(), win)
    assert np.array_equal(ex["reverse_index"].cpu().numpy(), O.reverse_index(win))
    assert np.array_equal(ex["cu_window_seqlens"].cpu().numpy(), cu_win)
    assert np.array_equal(ex["cu_seqlens"].cpu().numpy(), O.cu_seqlens(grid))
    assert np.array_equal(ex["pos_hw"].cpu().numpy(), O.pos_hw(grid))


def test_plan_random_grids_bit_exact():
    rng = np.random.default_rng(1)
    for _ in range(40):
        grid = [[int(rng.integers(1, 6)), int(rng.integers(1, 24)) * 2, int(rng.integers(1, 24)) * 2]
                for _ in range(int(rng.integers(1, 5)))]
        ex = _plan(grid).export()
        win, cu_win, _ = O.window_index(grid)
        assert np.array_equal(ex["window_index"].cpu().numpy(), win), grid
        assert np.array_equal(ex["cu_window_seqlens"].cpu().numpy(), cu_win), grid
        assert np.array_equal(ex["pos_hw"].cpu().numpy(), O.pos_hw(grid)), grid


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "indices_*.npz"))), ids=os.path.basename)
def test_plan_vs_hf_golden(path):
    g = np.load(path)
    ex = _plan(g["grid"].tolist()).export()
    assert np.array_equal(ex["window_index"].cpu().numpy(), g["window_index"])
    assert np.array_equal(ex["cu_window_seqlens"].cpu().numpy(), g["cu_dedup"])


def test_rotary_tables_close_to_oracle():
    grid = [[2, 10, 14], [1, 26, 46]]
    p = _plan(grid)
    cos, sin = p.export_rotary()
    win, _, _ = O.window_index(grid)
    rc, rs = O.rotary_tables(grid, win)
    assert np.abs(cos.cpu().numpy() - rc).max() < 1e-5 and np.abs(sin.cpu().numpy() - rs).max() < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_window_permute_and_inverse_are_bit_exact_gathers(dtype):
    """HF :478-484 permutes 4-row groups, :512-513 undoes it with argsort: gather(gather(x)) == x."""
    import video_r2_b200 as vr2
    grid = [[3, 26, 46], [2, 8, 8]]
    ex = _plan(grid).export()
    n_tok = ex["window_index"].numel()
    x = torch.randn(n_tok * 4, 1280, generator=torch.Generator().manual_seed(0)).to(dtype).cuda()
    grouped = x.reshape(n_tok, 4 * 1280)
    perm = vr2.gather_rows(grouped, ex["window_index"])
    assert torch.equal(perm, grouped[ex["window_index"]])
    back = vr2.gather_rows(perm, ex["reverse_index"])
    assert torch.equal(back, grouped)
    # full-size checksum-of-checksums property at config-2 scale
    big = _plan([[16, 32, 32]]).export()
    y = torch.randn(4096, 3584, generator=torch.Generator().manual_seed(1)).to(dtype).cuda()
    z = vr2.gather_rows(y, big["window_index"])
    assert torch.equal(z.view(torch.int16 if dtype == torch.bfloat16 else torch.int32).sum(0),
                       y.view(torch.int16 if dtype == torch.bfloat16 else torch.int32).sum(0))
    assert torch.equal(vr2.gather_rows(z, big["reverse_index"]), y)


def test_bad_grids_raise_value_error():
    import video_r2_b200 as vr2
    for bad in ([[1, 3, 4]], [[0, 4, 4]], [[1, 4, 5]]):
        with pytest.raises(ValueError):
            vr2.Plan(vr2.VisionConfig(), bad, "cuda:0")
