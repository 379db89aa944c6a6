"""ORACLE (test infrastructure, not product code): CPU restatement of the decoder-side M-RoPE of transformers
modeling_qwen2_5_vl.py — `Qwen2_5_VLRotaryEmbedding.forward` (freqs = inv_freq @ position_ids in fp32, emb = cat(freqs,
freqs), cos/sin * attention_scaling, cast to x.dtype) and `apply_multimodal_rotary_pos_emb` (chunks of mrope_section * 2
take the temporal / height / width rows cyclically; q * cos + rotate_half(q) * sin in the tensors' dtype).
Pinned by tests/test_oracle_mrope.py against the HF functions themselves (live, and golden vectors in tests/golden/).
Only tests/, smoke() and bench.py's CPU legs may import this module."""
import torch


def cos_sin(position_ids, inv_freq, attention_scaling=1.0, dtype=torch.float32):
    """position_ids [3, B, S] int64, inv_freq [D/2] fp32 -> cos, sin [3, B, S, D] of `dtype`."""
    freqs = position_ids[..., None].float() * inv_freq.float()[None, None, None, :]     # one fp32 product per element
    emb = torch.cat((freqs, freqs), dim=-1)
    return (emb.cos() * attention_scaling).to(dtype), (emb.sin() * attention_scaling).to(dtype)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_mrope(q, k, position_ids, inv_freq, mrope_section=(16, 24, 24), attention_scaling=1.0):
    """q [B, Hq, S, D], k [B, Hk, S, D] -> (q_embed, k_embed) in q's dtype, operation by operation as HF."""
    cos, sin = cos_sin(position_ids, inv_freq, attention_scaling, q.dtype)
    sec = list(mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1).unsqueeze(1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1).unsqueeze(1)
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def decoder_qkv(hidden, wq, bq, wk, bk, wv, bv, position_ids, inv_freq, mrope_section=(16, 24, 24), head_dim=128,
                dtype=torch.float32):
    """The first lines of `Qwen2_5_VLAttention.forward` (modeling_qwen2_5_vl.py:627-669): q/k/v = Linear(hidden) viewed
    as [B, heads, S, head_dim], multimodal rotary on q and k — every operation in `dtype` (float32: the reference the
    fused kernel is judged against; bfloat16: what the reference model itself computes, roundings included)."""
    import torch.nn.functional as F
    B, S, _ = hidden.shape
    x = hidden.to(dtype)
    q = F.linear(x, wq.to(dtype), bq.to(dtype)).view(B, S, -1, head_dim).transpose(1, 2)
    k = F.linear(x, wk.to(dtype), bk.to(dtype)).view(B, S, -1, head_dim).transpose(1, 2)
    v = F.linear(x, wv.to(dtype), bv.to(dtype)).view(B, S, -1, head_dim).transpose(1, 2)
    q, k = apply_mrope(q, k, position_ids, inv_freq, mrope_section)
    return q, k, v
