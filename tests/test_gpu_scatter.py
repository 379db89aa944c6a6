"""GPU: token positions, the row scatter and the tower forward with the fused scatter, bit-exact against the oracle
(and so against torch.masked_scatter, tests/test_oracle_scatter.py)."""
import numpy as np
import pytest
import torch

from oracle import scatter as S
from oracle import tower as T

pytestmark = pytest.mark.gpu
VID = 151656


@pytest.mark.parametrize("B,L,p", [(1, 7, 0.5), (3, 1000, 0.3), (2, 4099, 0.9), (5, 333, 0.0), (1, 70000, 0.6)])
def test_token_positions_bit_exact(B, L, p):
    import video_r2_b200 as vr2
    g = torch.Generator().manual_seed(B * 7 + L)
    ids = torch.randint(0, 1000, (B, L), generator=g)
    ids[torch.rand(B, L, generator=g) < p] = VID
    pos, count = vr2.token_positions(ids.cuda(), VID)
    ref = S.token_positions(ids.numpy(), VID)
    assert count == ref.shape[0]
    assert np.array_equal(pos.cpu().numpy().astype(np.int64), ref)
    # capacity smaller than the number of matches: the count is still the true count, positions are truncated
    pos2, count2 = vr2.token_positions(ids.cuda(), VID, capacity=3)
    assert count2 == count and np.array_equal(pos2.cpu().numpy().astype(np.int64), ref[:3])


@pytest.mark.parametrize("src_dt,dst_dt", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16),
                                           (torch.bfloat16, torch.float32), (torch.float32, torch.float32)])
def test_scatter_rows_matches_masked_scatter(src_dt, dst_dt):
    import video_r2_b200 as vr2
    g = torch.Generator().manual_seed(5)
    B, Sq, D = 3, 211, 3584
    ids = torch.randint(0, 1000, (B, Sq), generator=g)
    ids[torch.rand(B, Sq, generator=g) < 0.5] = VID
    n = int((ids == VID).sum())
    emb = torch.randn(B, Sq, D, generator=g).to(dst_dt)
    feats = torch.randn(n, D, generator=g).to(src_dt)
    ref = emb.masked_scatter((ids == VID).unsqueeze(-1).expand_as(emb), feats.to(dst_dt))
    dst = emb.clone().cuda()
    pos, count = vr2.token_positions(ids.cuda(), VID)
    assert count == n
    vr2.scatter_rows(feats.cuda(), dst, pos)
    assert torch.equal(dst.cpu(), ref)


def test_tower_forward_with_fused_scatter_equals_forward_then_masked_scatter():
    import video_r2_b200 as vr2
    cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), T.make_weights(cfg_o, seed=4), device="cuda:0")
    grid = [[2, 8, 12], [1, 6, 10]]
    n = sum(t * h * w for t, h, w in grid)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(n, 1176, generator=g).to(torch.bfloat16).cuda()
    rows = [[11, 12, 151652] + [VID] * 48 + [151653, 13, 14, 15, 16, 17],
            [21, 151652] + [VID] * 15 + [151653] + [22] * 39]
    ids = torch.tensor(rows)
    assert int((ids == VID).sum()) == n // 4
    for dt in (torch.bfloat16, torch.float32):
        emb = torch.randn(2, ids.shape[1], 3584, generator=g).to(dt).cuda()
        feats = tower(x, grid, out_dtype=dt)
        ref = emb.masked_scatter((ids.cuda() == VID).unsqueeze(-1).expand_as(emb), feats)
        out = vr2.embed_visual_tokens(tower, emb.clone(), ids.cuda(), VID, x, grid)
        assert torch.equal(out, ref)
    with pytest.raises(ValueError, match="features and video tokens do not match"):
        vr2.embed_visual_tokens(tower, emb.clone(), ids.cuda()[:, :-50], VID, x, grid)


def test_embedding_reuse_is_keyed_on_tensor_identity_and_version():
    import video_r2_b200 as vr2
    cfg_o = T.TowerConfig(depth=1, fullatt_block_indexes=(0,))
    tower = vr2.VisionTower(vr2.VisionConfig(depth=1, fullatt_block_indexes=(0,)), T.make_weights(cfg_o, seed=8), device="cuda:0")
    grid = [[2, 8, 12]]
    x = torch.randn(192, 1176, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16).cuda()
    a = tower(x, grid, reuse=2)
    b = tower(x, grid, reuse=2)
    assert b is a and tower.reuse_hits == 1              # same object, unmodified: served from the cache
    y = x.clone()
    c = tower(y, grid, reuse=2)
    assert c is not a and torch.equal(c, a)              # equal content, different tensor: recomputed
    x.mul_(2)                                            # in-place change bumps _version: no stale hit
    d = tower(x, grid, reuse=2)
    assert d is not a and not torch.equal(d, a) and tower.reuse_hits == 1
    assert torch.equal(tower(x, grid), d)                # and the uncached path agrees
