"""Per-token log-probabilities of the LM head without the [tokens, vocab] logits tensor (`vr2_lm_logprob`): what
Video-R2's GRPO computes as `selective_log_softmax(model(...).logits / temperature, input_ids)`
(train/src/trainer/grpo_trainer.py:723-741) from the decoder's last hidden states."""
from __future__ import annotations

import torch

from ._lib import check, lib
from .tower import _stream_ptr


def lm_logprob(hidden: torch.Tensor, weight: torch.Tensor, target: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
    """hidden [..., K] bf16, weight [V, K] bf16 (lm_head.weight), target [...] int64 -> fp32 [...]:
    log_softmax(hidden @ weight.T / temperature, -1).gather(-1, target).

    FORWARD ONLY.  It serves the passes of GRPO that run without a graph — the reference-policy and old-policy
    log-probabilities (grpo_trainer.py:813-833, under `torch.no_grad()` / `inference_mode`) and evaluation.  The
    policy pass that the loss back-propagates through (grpo_trainer.py:1094-1096) needs d logp / d hidden and
    d logp / d lm_head, which this kernel does not produce: calling it there would silently cut the gradient, so a
    call with grad enabled on an input that requires grad raises instead."""
    if torch.is_grad_enabled() and (hidden.requires_grad or weight.requires_grad):
        raise RuntimeError(
            "lm_logprob is forward-only (no backward kernel): call it under torch.no_grad() for the reference / "
            "old-policy passes, and keep `selective_log_softmax(model(...).logits / T, ids)` for the pass the loss "
            "differentiates")
    with torch.no_grad():
        return _lm_logprob(hidden, weight, target, temperature)


def _lm_logprob(hidden, weight, target, temperature):
    if not (hidden.is_cuda and weight.is_cuda) or hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("lm_logprob: hidden and weight must be bfloat16 CUDA tensors")
    K = hidden.shape[-1]
    if weight.dim() != 2 or weight.shape[1] != K:
        raise ValueError(f"lm_logprob: weight must be [V, {K}], got {tuple(weight.shape)}")
    if tuple(target.shape) != tuple(hidden.shape[:-1]):
        raise ValueError(f"lm_logprob: target shape {tuple(target.shape)} does not match hidden {tuple(hidden.shape[:-1])}")
    h = hidden.reshape(-1, K).contiguous()
    w = weight.contiguous()
    t = target.reshape(-1).to(device=h.device, dtype=torch.int64).contiguous()
    M, V = h.shape[0], w.shape[0]
    out = torch.empty(M, dtype=torch.float32, device=h.device)
    if M == 0:
        return out.reshape(target.shape)
    need = lib().vr2_lm_logprob_workspace_bytes(M, V)
    ws = torch.empty(need, dtype=torch.uint8, device=h.device)
    with torch.cuda.device(h.device):
        check(lib().vr2_lm_logprob(h.data_ptr(), w.data_ptr(), t.data_ptr(), M, V, K, float(temperature), out.data_ptr(),
                                   ws.data_ptr(), need, _stream_ptr(h.device)), "vr2_lm_logprob")
    return out.reshape(target.shape)
