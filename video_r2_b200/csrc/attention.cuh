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
                                int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16 = false,
                                bool reverse = false);   // reverse: query tiles from the last one down (GEMM_REVERSE_M)

}  // namespace vr2
