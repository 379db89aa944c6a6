"""GPU: vr2_mrope_apply against the oracle (and so against HF's rotary code).  The arithmetic after cos/sin is
reproduced operation by operation; cos/sin themselves come from CUDA's cosf/sinf instead of the CPU's, which may
differ in the last fp32 bit for large angles, so: fp32 within 2e-6 * (|q| + |k| scale), bf16 bit-identical on all but
a vanishing fraction of elements (those whose cos/sin sat within one fp32 ulp of a bf16 rounding boundary), never more
than one bf16 ulp apart."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mrope as O

pytestmark = pytest.mark.gpu


def _check(out, ref, dt):
    out, ref = out.float().cpu(), ref.float()
    if dt == torch.float32:
        assert (out - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    else:
        diff = (out - ref).abs()
        ulp = torch.maximum(ref.abs(), torch.tensor(2.0 ** -6)) * 2.0 ** -7      # one bf16 ulp at |ref| (>= a floor)
        assert (diff <= ulp).all(), (diff / ulp).max().item()
        assert (diff > 0).float().mean().item() < 2e-3


@pytest.mark.parametrize("name,dt", [("f32", torch.float32), ("bf16", torch.bfloat16)])
def test_mrope_matches_hf_golden(name, dt):
    import video_r2_b200 as vr2
    g = np.load(os.path.join(GOLDEN, "mrope_small.npz"))
    q, k = torch.from_numpy(g[f"q_{name}"]).to(dt).cuda(), torch.from_numpy(g[f"k_{name}"]).to(dt).cuda()
    rq, rk = vr2.apply_mrope(q, k, torch.from_numpy(g["pos"]).cuda(), torch.from_numpy(g["inv_freq"]).cuda(), g["section"].tolist())
    _check(rq, torch.from_numpy(g[f"rq_{name}"]), dt)
    _check(rk, torch.from_numpy(g[f"rk_{name}"]), dt)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_mrope_decoder_shape_and_positions_from_get_rope_index(dt):
    """Qwen2.5-VL-7B decoder geometry (28 q heads, 4 kv heads, head_dim 128) on position ids computed by this
    library's get_rope_index for a video prompt: the producer and the consumer of SURVEY row a12 together."""
    import video_r2_b200 as vr2
    ids = torch.tensor([[5, 6, 151652] + [151656] * 96 + [151653, 7, 8, 9], [151643] * 10 + [5, 6, 151652] + [151656] * 80 + [151653] + [7] * 9])
    mask = (ids != 151643).long()
    pos, _ = vr2.get_rope_index(ids.cuda(), None, torch.tensor([[4, 8, 12], [2, 8, 20]]).cuda(), torch.tensor([1.5, 0.5]).cuda(), mask.cuda())
    g = torch.Generator().manual_seed(3)
    B, S = ids.shape
    q = torch.randn(B, 28, S, 128, generator=g).to(dt)
    k = torch.randn(B, 4, S, 128, generator=g).to(dt)
    inv = 1.0 / (1000000.0 ** (torch.arange(0, 128, 2, dtype=torch.int64).float() / 128))
    rq, rk = vr2.apply_mrope(q.cuda(), k.cuda(), pos, inv.cuda())
    oq, ok = O.apply_mrope(q, k, pos.cpu(), inv)
    _check(rq, oq, dt)
    _check(rk, ok, dt)
    # in place
    q2, k2 = q.cuda().clone(), k.cuda().clone()
    vr2.apply_mrope(q2, k2, pos, inv.cuda(), inplace=True)
    assert torch.equal(q2, rq) and torch.equal(k2, rk)
    with pytest.raises(ValueError):
        vr2.apply_mrope(q.cuda(), k.cuda(), pos, inv.cuda(), mrope_section=(16, 24, 20))
