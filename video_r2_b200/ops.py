"""Single-kernel entry points of libvr2.so (the vr2_op_* functions of include/vr2.h), exposed
so tests can pin each kernel against the oracle.  Not a general operator library."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import VR2_BF16, VR2_F16, VR2_F32, check, lib


def _s(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def split_bf16(x: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi = bf16(x), lo = bf16(x - hi): the operand format of the "split" precision."""
    hi = x.to(torch.bfloat16)
    lo = (x.float() - hi.float()).to(torch.bfloat16)
    return hi, lo


def linear(a: torch.Tensor, w: torch.Tensor, bias=None, gelu: bool = False, out_dtype=torch.float32,
           a_lo: torch.Tensor = None) -> torch.Tensor:
    """a [M,K] bf16, w [N,K] bf16 -> act((a [+ a_lo]) @ w.T + bias), tcgen05 GEMM with fp32 accumulation.
    a and w may both be fp16 (the "f16" precision; kind::f16 rejects mixed formats); out_dtype float16 is then allowed."""
    assert a.is_cuda and a.dtype in (torch.bfloat16, torch.float16) and w.dtype == a.dtype
    assert out_dtype in (torch.float32, torch.bfloat16) or (out_dtype == torch.float16 and a.dtype == torch.float16)
    a, w = a.contiguous(), w.contiguous()
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    b = bias.to(torch.float32).contiguous() if bias is not None else None
    bp = b.data_ptr() if b is not None else None
    od = {torch.float32: VR2_F32, torch.bfloat16: VR2_BF16, torch.float16: VR2_F16}[out_dtype]
    with torch.cuda.device(a.device):
        if a.dtype == torch.float16:
            assert a_lo is None
            check(lib().vr2_op_linear_f16(a.data_ptr(), w.data_ptr(), bp, out.data_ptr(), od, M, N, K, int(gelu), _s(a)),
                  "vr2_op_linear_f16")
        elif a_lo is None:
            check(lib().vr2_op_linear(a.data_ptr(), w.data_ptr(), bp, out.data_ptr(), od, M, N, K, int(gelu), _s(a)),
                  "vr2_op_linear")
        else:
            assert a_lo.dtype == torch.bfloat16 and a_lo.shape == a.shape
            a_lo = a_lo.contiguous()
            check(lib().vr2_op_linear_split(a.data_ptr(), a_lo.data_ptr(), w.data_ptr(), bp, out.data_ptr(), od, M, N,
                                            K, int(gelu), _s(a)), "vr2_op_linear_split")
    return out


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """x fp32 [rows, dim] -> bf16(x * rsqrt(mean(x^2)+eps) * w)."""
    assert x.is_cuda and x.dtype == torch.float32
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(lib().vr2_op_rmsnorm(x.data_ptr(), w.to(torch.float32).contiguous().data_ptr(), out.data_ptr(),
                                   x.shape[0], x.shape[1], eps, _s(x)), "vr2_op_rmsnorm")
    return out


def attention(qkv: torch.Tensor, cu_seqlens, num_heads: int, qkv_lo: torch.Tensor = None):
    """qkv bf16 (or fp16) [N, 3*heads*80] (q | k | v), cu_seqlens int list/tensor -> same dtype [N, heads*80].
    With qkv_lo (the rounding residual of qkv) runs the split kernel and returns (out_hi, out_lo)."""
    assert qkv.is_cuda and qkv.dtype in (torch.bfloat16, torch.float16)
    qkv = qkv.contiguous()
    cu = [int(v) for v in (cu_seqlens.tolist() if isinstance(cu_seqlens, torch.Tensor) else cu_seqlens)]
    arr = (C.c_int32 * len(cu))(*cu)
    out = torch.empty((qkv.shape[0], num_heads * 80), dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(qkv.device):
        if qkv.dtype == torch.float16:
            assert qkv_lo is None
            check(lib().vr2_op_attention_f16(qkv.data_ptr(), out.data_ptr(), arr, len(cu) - 1, num_heads, _s(qkv)),
                  "vr2_op_attention_f16")
            return out
        if qkv_lo is None:
            check(lib().vr2_op_attention(qkv.data_ptr(), out.data_ptr(), arr, len(cu) - 1, num_heads, _s(qkv)),
                  "vr2_op_attention")
            return out
        assert qkv_lo.dtype == torch.bfloat16 and qkv_lo.shape == qkv.shape
        qkv_lo = qkv_lo.contiguous()
        out_lo = torch.empty_like(out)
        check(lib().vr2_op_attention_split(qkv.data_ptr(), qkv_lo.data_ptr(), out.data_ptr(), out_lo.data_ptr(), arr,
                                           len(cu) - 1, num_heads, _s(qkv)), "vr2_op_attention_split")
    return out, out_lo
