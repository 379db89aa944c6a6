// HBM-bound helpers: dtype casts, weight repacking at load time, and the one
// stand-alone RMSNorm the tower still needs (ln_q of the PatchMerger, HF
// modeling_qwen2_5_vl.py:57-71,:144-146 — the per-block norms are folded into the
// GEMM epilogues).  Vectorised 16-byte accesses, warp-shuffle reductions, grids
// sized in multiples of the SM count.
#include "elementwise.cuh"

#include <cuda_bf16.h>

#include "ptx.cuh"

namespace vr2 {

namespace {

// dst_lo (optional) receives the rounding residual bf16(v - bf16(v)) for the split ("precise") GEMMs
__global__ void cast_f32_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, uint2* __restrict__ dst_lo,
                                     int64_t n_vec) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(src + i);
    uint32_t h0, h1, l0, l1;
    split_bf16x2(v.x, v.y, h0, l0);
    split_bf16x2(v.z, v.w, h1, l1);
    dst[i] = make_uint2(h0, h1);
    if (dst_lo != nullptr) dst_lo[i] = make_uint2(l0, l1);
  }
}

// one warp per row; dim % 128 == 0 (32 lanes x float4)
__global__ void rmsnorm_bf16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo, int64_t rows,
                                    int dim, float eps, bool f16, unsigned int* __restrict__ sat) {
  const int lane = threadIdx.x & 31;
  float amax = 0.f;   // f16 range guard
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
    float ss = 0.f;
    for (int v = lane; v < dim / 4; v += 32) {
      const float4 a = xr[v];
      ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / static_cast<float>(dim) + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + row * dim);
    uint2* lrow = out_lo != nullptr ? reinterpret_cast<uint2*>(out_lo + row * dim) : nullptr;
    for (int v = lane; v < dim / 4; v += 32) {
      const float4 a = xr[v];
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + v);
      uint32_t h0, h1, l0 = 0, l1 = 0;
      if (f16) {                                    // fp16 activations (never with out_lo)
        const float v0 = a.x * rstd * g.x, v1 = a.y * rstd * g.y, v2 = a.z * rstd * g.z, v3 = a.w * rstd * g.w;
        h0 = pack_f16x2(v0, v1);
        h1 = pack_f16x2(v2, v3);
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1))), fmaxf(fabsf(v2), fabsf(v3)));
      } else {
        split_bf16x2(a.x * rstd * g.x, a.y * rstd * g.y, h0, l0);
        split_bf16x2(a.z * rstd * g.z, a.w * rstd * g.w, h1, l1);
      }
      orow[v] = make_uint2(h0, h1);
      if (lrow != nullptr) lrow[v] = make_uint2(l0, l1);
    }
  }
  if (f16 && sat != nullptr) sat_guard(amax, sat);
}

template <typename Src, typename Dst>
__global__ void repack_kernel(const Src* __restrict__ src, int64_t rows, int64_t cols, Dst* __restrict__ dst,
                              int64_t ld_dst, int64_t blk, int64_t blk_stride, int64_t offset) {
  const int64_t total = rows * cols;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    const int64_t dr = (r / blk) * blk_stride + offset + (r % blk);
    float v;
    if constexpr (sizeof(Src) == 2) v = __bfloat162float(src[e]); else v = static_cast<float>(src[e]);
    if constexpr (sizeof(Dst) == 2) dst[dr * ld_dst + c] = __float2bfloat16_rn(v); else dst[dr * ld_dst + c] = v;
  }
}

int grid_for(int64_t work_items, int threads) {
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

cudaError_t cast_f32_to_bf16(const float* src, __nv_bfloat16* dst, __nv_bfloat16* dst_lo, int64_t n,
                             cudaStream_t stream) {
  if (n % 4 != 0) return cudaErrorInvalidValue;
  cast_f32_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(src),
                                                                  reinterpret_cast<uint2*>(dst),
                                                                  reinterpret_cast<uint2*>(dst_lo), n / 4);
  return cudaGetLastError();
}

// bf16 -> fp16 (weights for the f16 activation mode: kind::f16 wants both operands in one format).  Exact for every
// bf16 value inside fp16's normal range; below 2^-14 the value is rounded to fp16's subnormal grid (|error| <= 2^-25).
__global__ void bf16_to_f16_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t pairs) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pairs;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t v = src[i];
    dst[i] = pack_f16x2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  }
}

cudaError_t cast_bf16_to_f16(const __nv_bfloat16* src, void* dst, int64_t n, cudaStream_t stream) {
  if (n % 2 != 0) return cudaErrorInvalidValue;
  bf16_to_f16_kernel<<<grid_for(n / 2, 256), 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(src),
                                                              static_cast<uint32_t*>(dst), n / 2);
  return cudaGetLastError();
}

cudaError_t rmsnorm_to_bf16(const float* x, const float* w, __nv_bfloat16* out, __nv_bfloat16* out_lo, int64_t rows,
                            int dim, float eps, cudaStream_t stream, bool f16, unsigned int* sat) {
  if (dim % 4 != 0 || (f16 && out_lo != nullptr)) return cudaErrorInvalidValue;
  rmsnorm_bf16_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(x, w, out, out_lo, rows, dim, eps, f16, sat);
  return cudaGetLastError();
}

cudaError_t repack_matrix(const void* src, bool src_bf16, int64_t rows, int64_t cols, void* dst, bool dst_bf16,
                          int64_t ld_dst, int64_t blk, int64_t blk_stride, int64_t offset, cudaStream_t stream) {
  const int g = grid_for(rows * cols, 256);
  if (src_bf16 && dst_bf16)
    repack_kernel<<<g, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), rows, cols, static_cast<__nv_bfloat16*>(dst), ld_dst, blk, blk_stride, offset);
  else if (src_bf16)
    repack_kernel<<<g, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), rows, cols, static_cast<float*>(dst), ld_dst, blk, blk_stride, offset);
  else if (dst_bf16)
    repack_kernel<<<g, 256, 0, stream>>>(static_cast<const float*>(src), rows, cols, static_cast<__nv_bfloat16*>(dst), ld_dst, blk, blk_stride, offset);
  else
    repack_kernel<<<g, 256, 0, stream>>>(static_cast<const float*>(src), rows, cols, static_cast<float*>(dst), ld_dst, blk, blk_stride, offset);
  return cudaGetLastError();
}

}  // namespace vr2
