"""CPU: the log-prob oracle against torch's own log_softmax + gather (the formula of TRL's selective_log_softmax that
grpo_trainer.py:739-741 calls) and the logsumexp identity; the reference's bf16 arithmetic stays within bf16 error of it."""
import torch

from oracle import logprob as L


def test_oracle_matches_torch_log_softmax_gather():
    g = torch.Generator().manual_seed(0)
    M, K, V = 37, 64, 1000
    h = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(V, K, generator=g) * 0.2).to(torch.bfloat16)
    t = torch.randint(0, V, (M,), generator=g)
    for temp in (1.0, 0.7):
        logits = h.double() @ w.double().T / temp
        ref = torch.log_softmax(logits, dim=-1).gather(-1, t[:, None]).squeeze(-1)      # selective_log_softmax, fp path
        ref2 = logits.gather(-1, t[:, None]).squeeze(-1) - torch.logsumexp(logits, -1)   # its logsumexp form
        out = L.lm_logprob(h, w, t, temp)
        assert torch.allclose(out, ref, atol=1e-12) and torch.allclose(out, ref2, atol=1e-12)
        assert (L.reference_bf16(h, w, t, temp).double() - out).abs().max() < 0.15    # bf16 logits of magnitude ~10
