"""GPU: vr2_lm_logprob against the float64 oracle on the same bf16-exact inputs.  fp32 accumulation of exact bf16
products and an fp32 log-sum-exp: tolerance 2e-3 absolute on log-probabilities of magnitude ~10 (the reference's own
bf16 arithmetic is ~0.05 away from the same yardstick, printed)."""
import pytest
import torch

from oracle import logprob as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,V,temp", [(200, 256, 1024, 1.0), (513, 3584, 5000, 0.7), (64, 3584, 152064, 1.0), (1, 64, 300, 2.0)])
def test_lm_logprob_matches_oracle(M, K, V, temp):
    import video_r2_b200 as vr2
    g = torch.Generator().manual_seed(M + V)
    h = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(V, K, generator=g) * (3.0 / K ** 0.5)).to(torch.bfloat16)
    t = torch.randint(0, V, (M,), generator=g)
    t[0] = V - 1
    t[-1] = 0
    out = vr2.lm_logprob(h.cuda(), w.cuda(), t.cuda(), temp)
    ref = L.lm_logprob(h, w, t, temp)
    err = (out.double().cpu() - ref).abs().max().item()
    err_ref = (L.reference_bf16(h[:16], w, t[:16], temp).double() - ref[:16]).abs().max().item()
    print(f"\\nM={M} K={K} V={V} T={temp}: max-abs {err:.2e} (reference bf16 arithmetic: {err_ref:.2e}), logp range "
          f"[{ref.min().item():.2f}, {ref.max().item():.2f}]")
    assert out.dtype == torch.float32 and err < 2e-3, err


def test_lm_logprob_shapes_and_errors():
    import video_r2_b200 as vr2
    g = torch.Generator().manual_seed(1)
    h = torch.randn(2, 9, 128, generator=g).to(torch.bfloat16).cuda()
    w = torch.randn(777, 128, generator=g).to(torch.bfloat16).cuda()
    t = torch.randint(0, 777, (2, 9), generator=g).cuda()
    out = vr2.lm_logprob(h, w, t)
    ref = L.lm_logprob(h.reshape(-1, 128).cpu(), w.cpu(), t.reshape(-1).cpu()).reshape(2, 9)
    assert out.shape == (2, 9) and (out.double().cpu() - ref).abs().max().item() < 2e-3
    bad = t.clone()
    bad[0, 0] = 777                                        # outside the vocabulary: NaN, the rest unaffected
    o2 = vr2.lm_logprob(h, w, bad)
    assert torch.isnan(o2[0, 0]) and torch.equal(o2.flatten()[1:], out.flatten()[1:])
    with pytest.raises(ValueError):
        vr2.lm_logprob(h.float(), w, t)
    with pytest.raises(ValueError):
        vr2.lm_logprob(h, w[:, :64], t)
