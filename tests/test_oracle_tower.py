"""CPU: pins oracle/tower.py against golden outputs of the HF vision module
(tests/golden/tower_*.npz, produced by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import tower as T

sys_path_ok = True


def load_case(path):
    g = np.load(path)
    cfg = T.TowerConfig(depth=int(g["depth"]), fullatt_block_indexes=tuple(int(v) for v in g["fullatt"]))
    w = T.make_weights(cfg, seed=int(g["weight_seed"]))
    grid = g["grid"].tolist()
    n = sum(t * h * ww for t, h, ww in grid)
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(int(g["pixel_seed"])))
    x = x.to(torch.bfloat16).to(torch.float32)
    return g, cfg, w, grid, x


def checksum(t):
    t = t.double().reshape(-1)
    return np.array([t.sum().item(), t.abs().sum().item(), t[:: max(1, t.numel() // 97)].sum().item()])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tower_*.npz"))), ids=os.path.basename)
def test_oracle_matches_hf_golden(path):
    g, cfg, w, grid, x = load_case(path)
    # the seeded inputs regenerate to the same numbers the fixture was made from
    assert np.allclose(checksum(x), g["pixel_checksum"], rtol=0, atol=1e-9)
    wsum = np.stack([checksum(w[k]) for k in sorted(w)][:8])
    assert np.allclose(wsum, g["weight_checksum"], rtol=0, atol=1e-9)
    out = T.tower_forward(w, cfg, x, grid).numpy()
    ref = g["out"]
    assert out.shape == ref.shape
    # two fp32 CPU evaluations of the same formula: only summation order differs
    assert np.abs(out - ref).max() < 2e-4, np.abs(out - ref).max()


def test_slice_independence():
    """A clip [t,h,w] equals t clips [1,h,w]: windows and 'full' attention never cross slices."""
    cfg = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    w = T.make_weights(cfg, seed=3)
    x = torch.randn(3 * 6 * 10, 1176, generator=torch.Generator().manual_seed(4))
    a = T.tower_forward(w, cfg, x, [[3, 6, 10]])
    b = T.tower_forward(w, cfg, x, [[1, 6, 10]] * 3)
    assert torch.equal(a, b)
