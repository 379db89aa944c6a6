"""Decoder-side M-RoPE application: `apply_multimodal_rotary_pos_emb(q, k, *rotary_emb(x, position_ids), mrope_section)`
of transformers modeling_qwen2_5_vl.py in one kernel (`vr2_mrope_apply`), with cos/sin formed from the position ids
that `get_rope_index` produced instead of being materialised as two [3, B, S, D] tensors per forward."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import VR2_BF16, VR2_F32, check, lib
from .tower import _stream_ptr


class _MropeFn(torch.autograd.Function):
    """The rotation is orthogonal per channel pair, so its vector-Jacobian product is the rotation by the NEGATED
    angle: the same kernel on the incoming gradients with negated position ids (cos is even, sin is odd — exact)."""

    @staticmethod
    def forward(ctx, q, k, position_ids, inv_freq, mrope_section, attention_scaling):
        ctx.save_for_backward(position_ids, inv_freq)
        ctx.sec, ctx.scaling = mrope_section, attention_scaling
        return _apply(q, k, position_ids, inv_freq, mrope_section, attention_scaling, False)

    @staticmethod
    def backward(ctx, gq, gk):
        position_ids, inv_freq = ctx.saved_tensors
        dq, dk = _apply(gq.contiguous(), gk.contiguous(), -position_ids, inv_freq, ctx.sec, ctx.scaling, False)
        return dq, dk, None, None, None, None


def apply_mrope(q: torch.Tensor, k: torch.Tensor, position_ids: torch.Tensor, inv_freq: torch.Tensor,
                mrope_section=(16, 24, 24), attention_scaling: float = 1.0, inplace: bool = False):
    """q [B, Hq, S, D], k [B, Hk, S, D] (bf16 or fp32, CUDA), position_ids [3, B, S] int64, inv_freq [D/2] fp32.
    -> (q_embed, k_embed), same shapes and dtype, equal to HF's result in that dtype.

    Differentiable with respect to q and k (training paths: `Qwen2_5_VLAttention.forward` with a trainable LLM or
    LoRA on q_proj / k_proj, train_grpo.sh `freeze_llm False`): under autograd the call goes through `_MropeFn`,
    whose backward is the inverse rotation by the same kernel.  `inplace=True` is refused when a gradient is needed."""
    if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad):
        if inplace:
            raise RuntimeError("apply_mrope(inplace=True) would overwrite tensors autograd needs; use inplace=False when training")
        return _MropeFn.apply(q, k, position_ids, inv_freq, tuple(int(v) for v in mrope_section), float(attention_scaling))
    return _apply(q, k, position_ids, inv_freq, mrope_section, attention_scaling, inplace)


@torch.no_grad()
def _apply(q, k, position_ids, inv_freq, mrope_section, attention_scaling, inplace):
    if not (q.is_cuda and k.is_cuda) or q.dtype != k.dtype or q.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("apply_mrope: q and k must be CUDA tensors of the same dtype (float32 or bfloat16)")
    if q.dim() != 4 or k.dim() != 4 or q.shape[0] != k.shape[0] or q.shape[2:] != k.shape[2:]:
        raise ValueError(f"apply_mrope: q {tuple(q.shape)} / k {tuple(k.shape)} must be [B, H, S, D] with equal B, S, D")
    B, Hq, S, D = q.shape
    Hk = k.shape[1]
    if tuple(position_ids.shape) != (3, B, S):
        raise ValueError(f"apply_mrope: position_ids must be [3, {B}, {S}], got {tuple(position_ids.shape)}")
    q, k = q.contiguous(), k.contiguous()
    pos = position_ids.to(device=q.device, dtype=torch.int64).contiguous()
    inv = inv_freq.to(device=q.device, dtype=torch.float32).contiguous()
    if inv.numel() != D // 2:
        raise ValueError(f"apply_mrope: inv_freq has {inv.numel()} entries for head_dim {D}")
    qo, ko = (q, k) if inplace else (torch.empty_like(q), torch.empty_like(k))
    sec = (C.c_int32 * 3)(*[int(v) for v in mrope_section])
    with torch.cuda.device(q.device):
        check(lib().vr2_mrope_apply(q.data_ptr(), k.data_ptr(), qo.data_ptr(), ko.data_ptr(),
                                    VR2_BF16 if q.dtype == torch.bfloat16 else VR2_F32, B, S, Hq, Hk, D, pos.data_ptr(),
                                    inv.data_ptr(), sec, float(attention_scaling), _stream_ptr(q.device)),
              "vr2_mrope_apply")
    return qo, ko
