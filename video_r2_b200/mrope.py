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


# ---------------------------------------------------------------------------------------------------------------
# The rotation fused into the decoder's q | k | v projection (vr2_mrope_table + vr2_decoder_qkv_mrope)
# ---------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def mrope_table(position_ids: torch.Tensor, inv_freq: torch.Tensor, mrope_section=(16, 24, 24), attention_scaling: float = 1.0):
    """position_ids [3, B, S] int64 (CUDA) -> (cos, sin) fp32 [B*S, D/2]: the rotary angles of every token with the
    mrope_section axis choice applied.  Computed once per decoder forward, shared by all layers."""
    if not position_ids.is_cuda or position_ids.dim() != 3 or position_ids.shape[0] != 3:
        raise ValueError("mrope_table: position_ids must be a CUDA tensor [3, B, S]")
    _, B, S = position_ids.shape
    pos = position_ids.to(torch.int64).contiguous()
    inv = inv_freq.to(device=pos.device, dtype=torch.float32).contiguous()
    half = inv.numel()
    cos = torch.empty((B * S, half), dtype=torch.float32, device=pos.device)
    sin = torch.empty_like(cos)
    sec = (C.c_int32 * 3)(*[int(v) for v in mrope_section])
    with torch.cuda.device(pos.device):
        check(lib().vr2_mrope_table(pos.data_ptr(), B, S, inv.data_ptr(), half, sec, float(attention_scaling), cos.data_ptr(),
                                    sin.data_ptr(), _stream_ptr(pos.device)), "vr2_mrope_table")
    return cos, sin


class FusedDecoderQKV:
    """q_proj / k_proj / v_proj of one `Qwen2_5_VLAttention` (modeling_qwen2_5_vl.py:627-669) as ONE tcgen05 GEMM whose
    epilogue adds the biases and applies the multimodal rotary embedding to the q and k heads.

        fused = vr2.FusedDecoderQKV(attn.q_proj, attn.k_proj, attn.v_proj, head_dim=128)      # once per layer
        cos, sin = vr2.mrope_table(position_ids, rotary_emb.inv_freq, mrope_section)           # once per forward
        q, k, v = fused(hidden_states, cos, sin)      # [B, Hq, S, 128], [B, Hkv, S, 128], [B, Hkv, S, 128] (strided views)

    FORWARD ONLY (prefill / generation, the no-grad reference and old-policy passes of GRPO): the weights are snapshotted
    (concatenated, bf16) at construction, and a call with grad enabled on an input that requires grad raises."""

    def __init__(self, q_proj, k_proj, v_proj, head_dim: int = 128):
        ws = [m.weight.detach() for m in (q_proj, k_proj, v_proj)]
        dev = ws[0].device
        if dev.type != "cuda":
            raise ValueError("FusedDecoderQKV: the projections must live on a CUDA device (no CPU path)")
        self.head_dim = head_dim
        self.nq, self.nk, self.nv = (w.shape[0] for w in ws)
        self.K = ws[0].shape[1]
        self.w = torch.cat([w.to(torch.bfloat16) for w in ws], dim=0).contiguous()
        bs = [(m.bias.detach().float() if m.bias is not None else torch.zeros(m.weight.shape[0], device=dev)) for m in (q_proj, k_proj, v_proj)]
        self.b = torch.cat(bs).contiguous()
        self.N = self.w.shape[0]
        if self.N % 256 or self.nq % head_dim or self.nk % head_dim or self.nv % head_dim:
            raise ValueError(f"FusedDecoderQKV: q/k/v widths {self.nq}/{self.nk}/{self.nv} must be multiples of head_dim "
                             f"{head_dim} and sum to a multiple of 256")

    def __call__(self, hidden_states: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
        if torch.is_grad_enabled() and hidden_states.requires_grad:
            raise RuntimeError("FusedDecoderQKV is forward-only: use it under torch.no_grad() (prefill, generation, the "
                               "reference / old-policy passes); keep the stock projections for the pass that is differentiated")
        with torch.no_grad():
            B, S, K = hidden_states.shape
            if K != self.K or hidden_states.dtype != torch.bfloat16:
                raise ValueError(f"FusedDecoderQKV: hidden_states must be bf16 [B, S, {self.K}]")
            if cos.shape != (B * S, self.head_dim // 2) or cos.dtype != torch.float32:
                raise ValueError(f"FusedDecoderQKV: cos / sin must be fp32 [{B * S}, {self.head_dim // 2}] (mrope_table)")
            h = hidden_states.contiguous().view(B * S, K)
            out = torch.empty((B * S, self.N), dtype=torch.bfloat16, device=h.device)
            with torch.cuda.device(h.device):
                check(lib().vr2_decoder_qkv_mrope(h.data_ptr(), self.w.data_ptr(), self.b.data_ptr(), B * S, self.N, K, self.head_dim,
                                                  self.nq + self.nk, cos.contiguous().data_ptr(), sin.contiguous().data_ptr(),
                                                  out.data_ptr(), _stream_ptr(h.device)), "vr2_decoder_qkv_mrope")
            D = self.head_dim
            q = out[:, :self.nq].view(B, S, self.nq // D, D).transpose(1, 2)
            k = out[:, self.nq:self.nq + self.nk].view(B, S, self.nk // D, D).transpose(1, 2)
            v = out[:, self.nq + self.nk:].view(B, S, self.nv // D, D).transpose(1, 2)
            return q, k, v
