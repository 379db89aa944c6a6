#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace vr2 {

// dst = bf16(src); dst_lo (nullable) = bf16(src - dst), the residual the split ("precise") GEMMs add back
cudaError_t cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, __nv_bfloat16* dst_lo, int64_t n,
                             cudaStream_t stream);

// out = bf16(x * rsqrt(mean(x^2) + eps) * w), one warp per row; out_lo (nullable) = its rounding residual
cudaError_t cast_bf16_to_f16(const __nv_bfloat16* src, void* dst_f16, int64_t n, cudaStream_t stream);
cudaError_t rmsnorm_to_bf16(const float* x, const float* w, __nv_bfloat16* out, __nv_bfloat16* out_lo, int64_t rows,
                            int dim, float eps, cudaStream_t stream, bool f16 = false, unsigned int* sat = nullptr);

// dst[(r / blk) * blk_stride + offset + r % blk][c] = src[r][c]  (dtype converted);
// dst must be zero-filled beforehand where padding is wanted.
cudaError_t repack_matrix(const void* src, bool src_bf16, int64_t rows, int64_t cols, void* dst, bool dst_bf16,
                          int64_t ld_dst, int64_t blk, int64_t blk_stride, int64_t offset, cudaStream_t stream);

}  // namespace vr2
