#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace vr2 {

// The split precision's attention (attention3.cu): operands are hi + lo bf16 pairs of the same [N, 3*hidden] layout,
// three tcgen05.mma products per contraction, output as a hi + lo pair.  seg as for attention_tc_launch.
cudaError_t attention_split_launch(const __nv_bfloat16* qkv, const __nv_bfloat16* qkv_lo, __nv_bfloat16* out,
                                   __nv_bfloat16* out_lo, const int2* seg, int n_rows, int num_heads, int64_t ld_qkv,
                                   int64_t ld_out, int num_sms, cudaStream_t stream);

// tcgen05 / TMEM / TMA kernel (attention2.cu).  seg[r] = (first, past-last) row of the segment row r belongs to
// (device, [n_rows]); query tiles are the consecutive 128-row chunks of [0, n_rows).
cudaError_t attention_tc_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* seg, int n_rows, int num_heads,
                                int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16 = false);

// Pipelined kernel for long segments (attention4.cu: 64-row key blocks, double-buffered scores / probabilities, the two
// streams of a CTA share the K|V blocks of one head).  Same arguments and results as attention_tc_launch.
cudaError_t attention_pipe_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* seg, int n_rows, int num_heads,
                                  int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16 = false);

// Which kernel serves a launch.  VR2_ATTN_KERNEL=tile|pipe forces one (A/B runs, tests); otherwise launches with long
// segments (the full-attention layers) take the pipelined kernel, the window layers the tile kernel.
enum AttnKind { ATTN_TILE = 0, ATTN_PIPE = 1 };
AttnKind attention_kind(bool long_segments);
cudaError_t attention_launch(AttnKind kind, const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* seg, int n_rows,
                             int num_heads, int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16);

}  // namespace vr2
