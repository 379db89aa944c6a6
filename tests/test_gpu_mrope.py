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


def test_fused_decoder_qkv_mrope_matches_fp32_and_beats_the_bf16_pipeline():
    """vr2_decoder_qkv_mrope (q|k|v projection + bias + M-RoPE in one tcgen05 GEMM) on the Qwen2.5-VL-7B decoder
    geometry (hidden 3584, 28 q heads, 4 kv heads, head_dim 128) with position ids from this library's get_rope_index:
    against the fp32 evaluation of the same lines of Qwen2_5_VLAttention.forward it must sit within bf16 output rounding,
    and closer than the reference's own all-bf16 pipeline does."""
    import video_r2_b200 as vr2
    ids = torch.tensor([[5, 6, 151652] + [151656] * 96 + [151653, 7, 8, 9], [151643] * 10 + [5, 6, 151652] + [151656] * 80 + [151653] + [7] * 9])
    mask = (ids != 151643).long()
    pos, _ = vr2.get_rope_index(ids.cuda(), None, torch.tensor([[4, 8, 12], [2, 8, 20]]).cuda(), torch.tensor([1.5, 0.5]).cuda(), mask.cuda())
    g = torch.Generator().manual_seed(11)
    B, S = ids.shape
    K, Hq, Hk, D = 3584, 28, 4, 128
    hidden = torch.randn(B, S, K, generator=g).to(torch.bfloat16)
    lin = {}
    for name, n in (("q", Hq * D), ("k", Hk * D), ("v", Hk * D)):
        m = torch.nn.Linear(K, n, bias=True)
        m.weight.data = (torch.randn(n, K, generator=g) * 0.03).to(torch.bfloat16)
        m.bias.data = torch.randn(n, generator=g).to(torch.bfloat16)
        lin[name] = m
    inv = 1.0 / (1000000.0 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    args = (hidden, lin["q"].weight.data, lin["q"].bias.data, lin["k"].weight.data, lin["k"].bias.data, lin["v"].weight.data,
            lin["v"].bias.data, pos.cpu(), inv)
    ref = O.decoder_qkv(*args, dtype=torch.float32)
    hf16 = O.decoder_qkv(*args, dtype=torch.bfloat16)

    fused = vr2.FusedDecoderQKV(lin["q"].cuda(), lin["k"].cuda(), lin["v"].cuda(), head_dim=D)
    cos, sin = vr2.mrope_table(pos, inv.cuda())
    # the table equals HF's cos/sin after the mrope_section selection (fp32, first half of the duplicated embedding)
    ocos, osin = O.cos_sin(pos.cpu(), inv)
    sec = [16, 24, 24]
    want_cos = torch.cat([m[i % 3] for i, m in enumerate(ocos[..., :64].split(sec, dim=-1))], dim=-1).reshape(B * S, 64)
    assert (cos.cpu() - want_cos).abs().max().item() < 2e-6
    out = fused(hidden.cuda(), cos, sin)
    for name, o, r, h in zip("qkv", out, ref, hf16):
        assert o.shape == r.shape and o.dtype == torch.bfloat16
        o = o.float().cpu()
        err = (o - r).abs()
        ulp = torch.maximum(r.abs(), torch.tensor(2.0 ** -6)) * 2.0 ** -8          # half a bf16 ulp at |ref| (>= a floor)
        frac_within = (err <= 1.01 * ulp + 1e-3).float().mean().item()
        e_ours, e_hf = err.max().item(), (h.float() - r).abs().max().item()
        print(f"\nfused decoder {name}: max-abs vs fp32 {e_ours:.4f} (HF bf16 pipeline {e_hf:.4f}); within output rounding {frac_within:.4f}")
        assert frac_within > 0.999 and e_ours <= e_hf + 1e-6, (name, e_ours, e_hf, frac_within)
    # forward only
    hg = hidden.cuda().clone().requires_grad_(True)
    with pytest.raises(RuntimeError):
        fused(hg, cos, sin)


def test_apply_mrope_is_differentiable():
    """Training paths reach the rotary through autograd: the backward is the rotation by the negated angle."""
    import video_r2_b200 as vr2
    g = torch.Generator().manual_seed(4)
    B, S, D = 2, 17, 128
    pos = torch.randint(0, 3000, (3, B, S), generator=g).cuda()
    inv = (1.0 / (1000000.0 ** (torch.arange(0, D, 2).float() / D))).cuda()
    q = torch.randn(B, 3, S, D, generator=g).cuda().requires_grad_(True)
    k = torch.randn(B, 2, S, D, generator=g).cuda().requires_grad_(True)
    rq, rk = vr2.apply_mrope(q, k, pos, inv)
    gq, gk = torch.randn_like(rq), torch.randn_like(rk)
    (rq * gq).sum().backward(retain_graph=True)
    (rk * gk).sum().backward()
    q2, k2 = q.detach().cpu().requires_grad_(True), k.detach().cpu().requires_grad_(True)
    oq, ok = O.apply_mrope(q2, k2, pos.cpu(), inv.cpu())
    ((oq * gq.cpu()).sum() + (ok * gk.cpu()).sum()).backward()
    assert (q.grad.cpu() - q2.grad).abs().max().item() < 1e-4 and (k.grad.cpu() - k2.grad).abs().max().item() < 1e-4
    with pytest.raises(RuntimeError):
        vr2.apply_mrope(q, k, pos, inv, inplace=True)
