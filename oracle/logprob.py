"""ORACLE (test infrastructure, not product code): CPU restatement of the per-token log-probabilities Video-R2's GRPO
takes from the LM head (train/src/trainer/grpo_trainer.py:723-741; TRL `selective_log_softmax`):

    logits = hidden @ lm_head.weight.T          # [tokens, vocab]
    logits = logits / temperature
    logp   = logits.log_softmax(-1).gather(-1, target[..., None]).squeeze(-1)

in float64 from the bf16-exact inputs (the reference itself rounds the logits to bf16 before the softmax; the kernel
keeps fp32 scores, so the exact value is the yardstick and `reference_bf16` below shows what the reference's own
arithmetic gives).  Pinned by tests/test_oracle_logprob.py against torch's log_softmax / TRL's formula.
Only tests/, smoke() and bench.py's CPU legs may import this module."""
import torch


def lm_logprob(hidden, weight, target, temperature=1.0):
    """hidden [M, K], weight [V, K], target [M] int64 -> float64 [M]."""
    logits = hidden.double() @ weight.double().T / temperature
    lse = torch.logsumexp(logits, dim=-1)
    return logits.gather(-1, target[:, None]).squeeze(-1) - lse


def reference_bf16(hidden, weight, target, temperature=1.0):
    """The reference's arithmetic: bf16 logits (fp32-accumulated matmul rounded to bf16), bf16 division, log_softmax of
    a bf16 row (computed in fp32 internally, rounded to bf16), gather."""
    logits = (hidden.float() @ weight.float().T).to(torch.bfloat16)
    logits = logits / temperature
    return torch.stack([torch.log_softmax(row, dim=-1)[t] for row, t in zip(logits, target)]).float()
