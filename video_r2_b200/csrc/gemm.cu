// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] * W[N,K]^T      A, W: bf16, K-major      accumulate: fp32 in TMEM
//
// One CTA per SM loops over 128 x BN output tiles (tile id -> n fastest, so the
// 148 resident CTAs share a band of A rows and all of W through L2).  Roles:
//   warp 4 (one lane)  TMA producer: 128B-swizzled A/W k-blocks into a 4-stage ring
//   warp 5 (one lane)  tcgen05.mma issuer: 4 x (128 x BN x 16) per k-block, two
//                      accumulator stages in TMEM (2 x 256 columns)
//   warps 0-3          epilogue: tcgen05.ld of the finished accumulator (thread =
//                      row), fused bias / residual / RMS scale / rotary / SwiGLU /
//                      GELU, stores; overlaps the next tile's MMAs
// Replaces the cuBLAS/cuDNN calls behind nn.Linear / nn.Conv3d in
// HF modeling_qwen2_5_vl.py:108-114 (patch embed), :221-287 (qkv, proj),
// :87-88 (gate/up/down), :144-146 (merger).
#include "gemm.cuh"

#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "ptx.cuh"

namespace vr2 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;   // one 128-byte swizzle atom of bf16
constexpr int UK = 16;   // K per tcgen05.mma (bf16)
constexpr int kThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages

template <int BN>
struct Cfg {
  static constexpr int a_bytes = BM * BK * 2;
  static constexpr int b_bytes = BN * BK * 2;
  static constexpr int stage_bytes = a_bytes + b_bytes;
  static constexpr int stages = (196 * 1024) / stage_bytes;
  static constexpr int bar_bytes = 256;
  static constexpr int smem_bytes = stages * stage_bytes + bar_bytes + 1024;  // +1024: manual alignment
  static_assert(stage_bytes % 1024 == 0, "stage must keep 1024-byte alignment for SWIZZLE_128B");
  static_assert(stages >= 3, "pipeline too shallow");
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float v) { return v / (1.0f + __expf(-v)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::stages * C::stage_bytes);
  uint64_t* empty_bar = full_bar + C::stages;
  uint64_t* acc_full = empty_bar + C::stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_n_blocks = (p.N + BN - 1) / BN;
  const int num_m_blocks = (p.M + BM - 1) / BM;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_k_blocks = (p.K + BK - 1) / BK;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n_blocks) * BM;
        const int n0 = (tile % num_n_blocks) * BN;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const uint32_t s = it % C::stages;
          const uint32_t ph = (it / C::stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = smem + s * C::stage_bytes;
          uint8_t* sb = sa + C::a_bytes;
          mbar_expect_tx(&full_bar[s], C::stage_bytes);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      uint32_t it = 0;
      uint32_t local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const uint32_t as = local & 1u;
        const uint32_t use = local >> 1;
        mbar_wait(&acc_empty[as], (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccStride;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const uint32_t s = it % C::stages;
          const uint32_t ph = (it / C::stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * C::stage_bytes);
          const uint32_t b_addr = a_addr + C::a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * UK * 2);
            const uint64_t db = umma_desc_sw128(b_addr + k * UK * 2);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
        }
        umma_commit(&acc_full[as]);    // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0-3)
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int n_blk = tile % num_n_blocks;
      const int m0 = (tile / num_n_blocks) * BM;
      const int n0 = n_blk * BN;
      const uint32_t as = local & 1u;
      const uint32_t use = local >> 1;
      mbar_wait(&acc_full[as], use & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + as * kAccStride;
      const int r = m0 + warp * 32 + lane;
      const bool valid = r < p.M;

      if constexpr (EPI == EPI_STORE) {
        const int64_t dest = valid ? (p.row_map ? p.row_map[r] : r) : 0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (valid) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = n0 + c + j;
              float t = __uint_as_float(v[j]);
              if (p.bias != nullptr && col < p.N) t += __ldg(p.bias + col);
              if (p.flags & GEMM_GELU) t = gelu_erf(t);
              f[j] = t;
            }
            if (p.flags & GEMM_OUT_F32) {
              float* o = p.out_f32 + dest * p.ld_out + n0 + c;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (n0 + c + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              __nv_bfloat16* o = p.out_bf16 + dest * p.ld_out + n0 + c;
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                if (n0 + c + j < p.N)
                  *reinterpret_cast<uint4*>(o + j) =
                      make_uint4(pack_bf16x2(f[j], f[j + 1]), pack_bf16x2(f[j + 2], f[j + 3]),
                                 pack_bf16x2(f[j + 4], f[j + 5]), pack_bf16x2(f[j + 6], f[j + 7]));
            }
          }
        }
      } else if constexpr (EPI == EPI_RESID) {
        const int64_t dest = valid ? (p.row_map ? p.row_map[r] : r) : 0;
        float ss = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          tmem_ld_wait();
          if (valid) {
            float* xr = p.x + dest * p.ldx + n0 + c;
            __nv_bfloat16* xb = p.out_bf16 + dest * p.ld_out + n0 + c;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float f[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) f[q] = __uint_as_float(v[j + q]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              if (p.flags & GEMM_ADD_RESID) {
                const float4 x0 = *reinterpret_cast<const float4*>(xr + j);
                const float4 x1 = *reinterpret_cast<const float4*>(xr + j + 4);
                f[0] += x0.x; f[1] += x0.y; f[2] += x0.z; f[3] += x0.w;
                f[4] += x1.x; f[5] += x1.y; f[6] += x1.z; f[7] += x1.w;
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) ss = fmaf(f[q], f[q], ss);
              *reinterpret_cast<float4*>(xr + j) = make_float4(f[0], f[1], f[2], f[3]);
              *reinterpret_cast<float4*>(xr + j + 4) = make_float4(f[4], f[5], f[6], f[7]);
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.norm_w + n0 + c + j));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.norm_w + n0 + c + j + 4));
              *reinterpret_cast<uint4*>(xb + j) =
                  make_uint4(pack_bf16x2(f[0] * w0.x, f[1] * w0.y), pack_bf16x2(f[2] * w0.z, f[3] * w0.w),
                             pack_bf16x2(f[4] * w1.x, f[5] * w1.y), pack_bf16x2(f[6] * w1.z, f[7] * w1.w));
            }
          }
        }
        if (valid) p.ssq[dest * num_n_blocks + n_blk] = ss;
      } else if constexpr (EPI == EPI_QKV) {
        // Tile = 3 heads of 80; rotary pairs (d, d+40) live in the same thread's row.
        float rstd = 0.f;
        if (valid) {
          float s = 0.f;
          for (int q = 0; q < p.ssq_parts; ++q) s += p.ssq_in[static_cast<int64_t>(r) * p.ssq_parts + q];
          rstd = rsqrtf(s * p.inv_dim + p.eps);
        }
#pragma unroll 1
        for (int hd = 0; hd < BN / 80; ++hd) {
          const int col0 = n0 + hd * 80;
          const bool rot = col0 < 2 * (p.N / 3);  // q and k heads rotate, v heads do not
#pragma unroll 1
          for (int i0 = 0; i0 < 40; i0 += 8) {
            uint32_t lo[8], hi[8];
            tmem_ld8(taddr + hd * 80 + i0, lo);
            tmem_ld8(taddr + hd * 80 + 40 + i0, hi);
            tmem_ld_wait();
            if (valid) {
              float a[8], b[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                a[q] = fmaf(__uint_as_float(lo[q]), rstd, __ldg(p.bias + col0 + i0 + q));
                b[q] = fmaf(__uint_as_float(hi[q]), rstd, __ldg(p.bias + col0 + 40 + i0 + q));
              }
              if (rot) {
                const float* cr = p.cos_tab + static_cast<int64_t>(r) * 40 + i0;
                const float* sr = p.sin_tab + static_cast<int64_t>(r) * 40 + i0;
                const float4 c0 = *reinterpret_cast<const float4*>(cr);
                const float4 c1 = *reinterpret_cast<const float4*>(cr + 4);
                const float4 s0 = *reinterpret_cast<const float4*>(sr);
                const float4 s1 = *reinterpret_cast<const float4*>(sr + 4);
                const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  // q*cos + rotate_half(q)*sin, fp32 (HF :161-166), no FMA contraction surprises needed
                  const float na = a[q] * cs[q] - b[q] * sn[q];
                  const float nb = b[q] * cs[q] + a[q] * sn[q];
                  a[q] = na;
                  b[q] = nb;
                }
              }
              __nv_bfloat16* o = p.out_bf16 + static_cast<int64_t>(r) * p.ld_out + col0 + i0;
              *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]),
                                                        pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
              *reinterpret_cast<uint4*>(o + 40) = make_uint4(pack_bf16x2(b[0], b[1]), pack_bf16x2(b[2], b[3]),
                                                             pack_bf16x2(b[4], b[5]), pack_bf16x2(b[6], b[7]));
            }
          }
        }
      } else if constexpr (EPI == EPI_SWIGLU) {
        // Tile columns [0,128) are gate rows, [128,256) the matching up rows.
        float rstd = 0.f;
        if (valid) {
          float s = 0.f;
          for (int q = 0; q < p.ssq_parts; ++q) s += p.ssq_in[static_cast<int64_t>(r) * p.ssq_parts + q];
          rstd = rsqrtf(s * p.inv_dim + p.eps);
        }
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 16) {
          uint32_t g[16], u[16];
          tmem_ld16(taddr + c, g);
          tmem_ld16(taddr + BN / 2 + c, u);
          tmem_ld_wait();
          if (valid) {
            float h[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float gv = fmaf(__uint_as_float(g[q]), rstd, __ldg(p.bias + n0 + c + q));
              const float uv = fmaf(__uint_as_float(u[q]), rstd, __ldg(p.bias + n0 + BN / 2 + c + q));
              h[q] = silu(gv) * uv;
            }
            __nv_bfloat16* o = p.out_bf16 + static_cast<int64_t>(r) * p.ld_out + n_blk * (BN / 2) + c;
            *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]),
                                                      pack_bf16x2(h[4], h[5]), pack_bf16x2(h[6], h[7]));
            *reinterpret_cast<uint4*>(o + 8) = make_uint4(pack_bf16x2(h[8], h[9]), pack_bf16x2(h[10], h[11]),
                                                          pack_bf16x2(h[12], h[13]), pack_bf16x2(h[14], h[15]));
          }
        }
      }
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn resolve_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  });
  return fn;
}

template <int BN, int EPI>
cudaError_t launch(const GemmArgs& a, int num_sms, cudaStream_t stream, const char** err) {
  using C = Cfg<BN>;
  CUtensorMap tmA, tmB;
  if (!tma_encode_2d_bf16(&tmA, a.A, a.K, a.M, a.lda * 2, BK, BM) ||
      !tma_encode_2d_bf16(&tmB, a.W, a.K, a.N, a.ldw * 2, BK, BN)) {
    *err = "cuTensorMapEncodeTiled failed (alignment or driver entry point)";
    return cudaErrorInvalidValue;
  }
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_kernel<BN, EPI><<<grid, kThreads, C::smem_bytes, stream>>>(tmA, tmB, a);
  return cudaGetLastError();
}

}  // namespace

bool tma_encode_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_rows) {
  EncodeFn fn = resolve_encode();
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

cudaError_t gemm_launch(const GemmArgs& a, int num_sms, cudaStream_t stream, const char** err) {
  static const char* none = "";
  *err = none;
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) {
    *err = "gemm: empty problem";
    return cudaErrorInvalidValue;
  }
  if ((a.lda * 2) % 16 != 0 || (a.ldw * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
      (reinterpret_cast<uintptr_t>(a.W) & 15)) {
    *err = "gemm: operand rows must be 16-byte aligned for TMA";
    return cudaErrorInvalidValue;
  }
  switch (a.epilogue) {
    case EPI_STORE:
      if ((a.flags & GEMM_OUT_F32) ? a.out_f32 == nullptr : a.out_bf16 == nullptr) {
        *err = "gemm: missing output";
        return cudaErrorInvalidValue;
      }
      if (a.N % 8 != 0) {
        *err = "gemm: EPI_STORE needs N % 8 == 0";
        return cudaErrorInvalidValue;
      }
      return launch<256, EPI_STORE>(a, num_sms, stream, err);
    case EPI_RESID:
      if (a.N % 256 != 0 || a.x == nullptr || a.out_bf16 == nullptr || a.norm_w == nullptr || a.ssq == nullptr) {
        *err = "gemm: EPI_RESID needs N % 256 == 0 and x/xb/norm_w/ssq";
        return cudaErrorInvalidValue;
      }
      return launch<256, EPI_RESID>(a, num_sms, stream, err);
    case EPI_QKV:
      if (a.N % 240 != 0 || a.bias == nullptr || a.ssq_in == nullptr || a.cos_tab == nullptr ||
          a.sin_tab == nullptr || a.out_bf16 == nullptr) {
        *err = "gemm: EPI_QKV needs N % 240 == 0 and bias/ssq/cos/sin/out";
        return cudaErrorInvalidValue;
      }
      return launch<240, EPI_QKV>(a, num_sms, stream, err);
    case EPI_SWIGLU:
      if (a.N % 256 != 0 || a.bias == nullptr || a.ssq_in == nullptr || a.out_bf16 == nullptr) {
        *err = "gemm: EPI_SWIGLU needs N % 256 == 0 and bias/ssq/out";
        return cudaErrorInvalidValue;
      }
      return launch<256, EPI_SWIGLU>(a, num_sms, stream, err);
    default:
      *err = "gemm: unknown epilogue";
      return cudaErrorInvalidValue;
  }
}

}  // namespace vr2
