"""GPU: the full tower through the C ABI against the fp32 oracle and the HF golden outputs.

Tolerance (BASELINE.json north_star): bf16 embeddings within max-abs 2e-2 and cosine >= 0.9995 of
the fp32 reference.  Two operand precisions exist (include/vr2.h vr2_precision):
  * "split" (hi + lo bf16 operands): asserted at max-abs <= 2e-2 AND cosine >= 0.9995 on every golden,
    including the depth-32 ragged one, with bf16 output — the north-star bound, met.
  * "f16" (fp16 activations x bf16 weights, the speed of "bf16"): the same assertion, met.
  * "bf16" (the arithmetic HF's own bf16 path uses; the throughput mode): cosine >= 0.9995 asserted; max-abs
    is bounded by the operand-rounding floor SURVEY §7.3 measured for ANY bf16-operand implementation of a
    32-block tower on this init (0.06; HF's own bf16 path: 0.165), asserted at 8e-2 and printed."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import tower as T

pytestmark = pytest.mark.gpu


def _stats(out, ref):
    out, ref = out.float().cpu(), ref.float().cpu()
    err = (out - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(out.flatten(), ref.flatten(), dim=0).item()
    row_cos = torch.nn.functional.cosine_similarity(out, ref, dim=1).min().item()
    return err, cos, row_cos


def _tower(cfg_o, weights, precision="bf16"):
    import video_r2_b200 as vr2
    cfg = vr2.VisionConfig(depth=cfg_o.depth, fullatt_block_indexes=tuple(cfg_o.fullatt_block_indexes))
    return vr2.VisionTower(cfg, weights, device="cuda:0", precision=precision)


TOL_MAX_ABS = 2e-2   # BASELINE.json north_star
TOL_COSINE = 0.9995


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tower_*.npz"))), ids=os.path.basename)
def test_tower_split_precision_meets_north_star_tolerance(path):
    g = np.load(path)
    cfg = T.TowerConfig(depth=int(g["depth"]), fullatt_block_indexes=tuple(int(v) for v in g["fullatt"]))
    w = T.make_weights(cfg, seed=int(g["weight_seed"]))
    grid = g["grid"].tolist()
    n = sum(t * h * ww for t, h, ww in grid)
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(int(g["pixel_seed"]))).to(torch.bfloat16)
    tower = _tower(cfg, w, precision="split")
    ref = torch.from_numpy(g["out"])
    out32 = tower(x.cuda(), grid, out_dtype=torch.float32)
    out16 = tower(x.cuda(), grid)
    assert out16.dtype == torch.bfloat16
    e32, c32, r32 = _stats(out32, ref)
    e16, c16, r16 = _stats(out16, ref)
    print(f"\n{os.path.basename(path)} [split]: fp32-out max-abs {e32:.5f} cos {c32:.7f} | bf16-out max-abs {e16:.5f} "
          f"cos {c16:.7f} min-row-cos {r16:.7f}")
    assert e16 <= TOL_MAX_ABS and c16 >= TOL_COSINE and r16 >= TOL_COSINE, (e16, c16, r16)
    assert e32 <= 5e-3, e32   # without the final bf16 rounding of a +-7 output (0.0156) the path is ~1e-3 or better
    # fp32 pixels take the split cast (hi + lo) path: same result for bf16-exact pixels
    assert torch.equal(tower(x.float().cuda(), grid), out32)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tower_*.npz"))), ids=os.path.basename)
def test_tower_f16_precision_meets_north_star_tolerance(path):
    """fp16 activations x bf16 weights: the speed of the bf16 mode, inside the north-star bound."""
    g = np.load(path)
    cfg = T.TowerConfig(depth=int(g["depth"]), fullatt_block_indexes=tuple(int(v) for v in g["fullatt"]))
    w = T.make_weights(cfg, seed=int(g["weight_seed"]))
    grid = g["grid"].tolist()
    n = sum(t * h * ww for t, h, ww in grid)
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(int(g["pixel_seed"]))).to(torch.bfloat16)
    tower = _tower(cfg, w, precision="f16")
    ref = torch.from_numpy(g["out"])
    out32 = tower(x.cuda(), grid, out_dtype=torch.float32)
    out16 = tower(x.cuda(), grid)
    assert out16.dtype == torch.bfloat16
    e32, c32, r32 = _stats(out32, ref)
    e16, c16, r16 = _stats(out16, ref)
    print(f"\n{os.path.basename(path)} [f16]: fp32-out max-abs {e32:.5f} cos {c32:.7f} | bf16-out max-abs {e16:.5f} "
          f"cos {c16:.7f} min-row-cos {r16:.7f}")
    assert e16 <= TOL_MAX_ABS and c16 >= TOL_COSINE and r16 >= TOL_COSINE, (e16, c16, r16)
    assert e32 <= 1.2e-2, e32
    assert torch.equal(tower(x.cuda(), grid, out_dtype=torch.float32), out32)      # deterministic


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tower_*.npz"))), ids=os.path.basename)
def test_tower_vs_hf_golden(path):
    g = np.load(path)
    cfg = T.TowerConfig(depth=int(g["depth"]), fullatt_block_indexes=tuple(int(v) for v in g["fullatt"]))
    w = T.make_weights(cfg, seed=int(g["weight_seed"]))
    grid = g["grid"].tolist()
    n = sum(t * h * ww for t, h, ww in grid)
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(int(g["pixel_seed"]))).to(torch.bfloat16)
    tower = _tower(cfg, w)
    ref = torch.from_numpy(g["out"])
    out32 = tower(x.cuda(), torch.tensor(grid).cuda(), out_dtype=torch.float32)
    out16 = tower(x.cuda(), grid)  # bf16 in -> bf16 out, the contract callers see
    assert out16.dtype == torch.bfloat16 and out16.shape == ref.shape
    e32, c32, r32 = _stats(out32, ref)
    e16, c16, r16 = _stats(out16, ref)
    print(f"\n{os.path.basename(path)}: fp32-out max-abs {e32:.4f} cos {c32:.6f} min-row-cos {r32:.6f} | "
          f"bf16-out max-abs {e16:.4f} cos {c16:.6f}")
    assert c16 >= 0.9995 and r16 >= 0.9995
    # bf16-operand floor measured in SURVEY §7.3: merger alone 0.0205, whole depth-32 tower 0.06
    bound = 3e-2 if cfg.depth <= 4 else 8e-2
    assert e32 <= bound, (e32, bound)


def test_tower_vs_oracle_multi_clip_and_dummy():
    cfg = T.TowerConfig(depth=3, fullatt_block_indexes=(1,))
    w = T.make_weights(cfg, seed=7)
    tower = _tower(cfg, w)
    for grid, seed in (([[2, 10, 14], [1, 4, 6], [1, 26, 6]], 8), ([[1, 28, 28]], 9)):
        n = sum(t * h * ww for t, h, ww in grid)
        x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16)
        if grid == [[1, 28, 28]]:
            x = torch.zeros_like(x)  # the ZeRO-3 dummy of monkey_patch_forward.py:362-366
        ref = T.tower_forward(w, cfg, x.float(), grid)
        out = tower(x.cuda(), grid, out_dtype=torch.float32)
        e, c, r = _stats(out, ref)
        print(f"\n{grid}: max-abs {e:.4f} cos {c:.6f}")
        assert e <= 3e-2 and c >= 0.9995, (grid, e, c)
    # fp32 pixels in -> fp32 out (the oracle's own dtype), identical numbers to the bf16 path for bf16-exact pixels
    grid = [[2, 10, 14]]
    x = torch.randn(280, 1176, generator=torch.Generator().manual_seed(10)).to(torch.bfloat16)
    a = tower(x.cuda(), grid, out_dtype=torch.float32)
    b = tower(x.float().cuda(), grid)
    assert b.dtype == torch.float32 and torch.equal(a, b)


def test_slice_independence_and_determinism():
    """[t,h,w] == t x [1,h,w] bit for bit (the sharding unit), and repeat runs are identical."""
    cfg = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    w = T.make_weights(cfg, seed=3)
    tower = _tower(cfg, w)
    x = torch.randn(4 * 12 * 20, 1176, generator=torch.Generator().manual_seed(4)).to(torch.bfloat16).cuda()
    a = tower(x, [[4, 12, 20]]).clone()
    b = tower(x, [[1, 12, 20]] * 4).clone()
    c = tower(x, [[4, 12, 20]]).clone()
    assert torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.parametrize("prec", ["f16", "bf16"])
def test_traversal_order_does_not_change_a_bit(prec, monkeypatch):
    """Consecutive launches walk the row blocks in opposite directions (GEMM_REVERSE_M / attention `reverse`)
    so that each starts on the rows its producer left in L2: a scheduling choice — the outputs are the same bits as with
    every launch ascending (VR2_ALTERNATE=0), on a grid with partial windows and a ragged last tile."""
    import video_r2_b200 as vr2
    cfg = T.TowerConfig(depth=4, fullatt_block_indexes=(1, 3))
    w = T.make_weights(cfg, seed=5)
    tower = vr2.VisionTower(vr2.VisionConfig(depth=4, fullatt_block_indexes=(1, 3)), w, device="cuda:0", precision=prec)
    grid = [[3, 26, 18], [2, 32, 32]]
    n = sum(t * h * ww for t, h, ww in grid)
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16).cuda()
    monkeypatch.setenv("VR2_ALTERNATE", "1")
    a = tower(x, grid, out_dtype=torch.float32).clone()
    monkeypatch.setenv("VR2_ALTERNATE", "0")
    b = tower(x, grid, out_dtype=torch.float32).clone()
    assert torch.equal(a, b)


def test_shape_errors_raise_value_error():
    cfg = T.TowerConfig(depth=1, fullatt_block_indexes=(0,))
    tower = _tower(cfg, T.make_weights(cfg, seed=1))
    with pytest.raises(ValueError):
        tower(torch.zeros(10, 1176, device="cuda"), [[1, 4, 4]])
    with pytest.raises(ValueError):
        tower(torch.zeros(16, 1000, device="cuda"), [[1, 4, 4]])


def test_tall_frame_uses_per_row_rotary_tables():
    """More than 64 patch rows (frames taller than 896 px): the CTA-pair QKV kernel's 64-position table does not
    apply and the single-CTA kernel with per-row tables runs instead — same numbers."""
    cfg = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    w = T.make_weights(cfg, seed=12)
    tower = _tower(cfg, w)
    for grid in ([[1, 66, 6]], [[2, 64, 8]], [[1, 4, 70], [1, 8, 8]]):
        n = sum(t * h * ww for t, h, ww in grid)
        x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(13)).to(torch.bfloat16)
        ref = T.tower_forward(w, cfg, x.float(), grid)
        out = tower(x.cuda(), grid, out_dtype=torch.float32)
        e, c, r = _stats(out, ref)
        print(f"\n{grid}: max-abs {e:.4f} cos {c:.6f}")
        assert e <= 3e-2 and c >= 0.9995, (grid, e, c)


def test_grid_tensor_is_recognised_without_a_second_readback():
    """The same `grid_thw` tensor object, unmodified, maps to its plan without another D2H; an in-place change of the
    tensor (version bump) is seen."""
    cfg = T.TowerConfig(depth=1, fullatt_block_indexes=(0,))
    tower = _tower(cfg, T.make_weights(cfg, seed=2))
    g = torch.tensor([[1, 4, 6]], device="cuda")
    x = torch.randn(48, 1176, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
    a = tower(torch.cat([x[:24], x[:24]]), g.clone().mul_(torch.tensor([2, 1, 1], device="cuda"))).clone()   # [[2,4,6]]
    p1 = tower.plan(g)
    assert tower.plan(g) is p1 and id(g) in tower._grid_keys
    b = tower(x[:24], g)
    g[0, 0] = 2                                    # same object, new contents
    p2 = tower.plan(g)
    assert p2 is not p1 and p2.grid == ((2, 4, 6),)
    assert torch.equal(tower(torch.cat([x[:24], x[:24]]), g), a) and b.shape[0] == 6
