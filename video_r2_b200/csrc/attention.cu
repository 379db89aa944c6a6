// Variable-length, non-causal attention over cu_seqlens segments for the vision
// tower (HF modeling_qwen2_5_vl.py:221-287): window layers run over
// cu_window_seqlens (<= 64 patches per window), layers 7/15/23/31 over one
// temporal slice (h*w patches).  16 heads x head_dim 80, q/k/v bf16 in one
// [N, 3840] buffer written by the QKV GEMM epilogue (rotary already applied).
//
// v1 kernel: flash-style online softmax, one CTA = one 64-row query tile of one
// head, 4 warps x 16 rows, K/V streamed through shared memory in 64-row tiles,
// mma.sync.m16n8k16 bf16 with fp32 accumulation and fp32 softmax.  The tile
// schedule (query tile -> its segment) is precomputed by the plan.
#include "attention.cuh"

#include <cuda_bf16.h>

#include "ptx.cuh"

namespace vr2 {

namespace {

constexpr int HD = 80;    // head dim
constexpr int HDP = 88;   // padded smem row (176 B: conflict-free ldmatrix)
constexpr int TQ = 64;
constexpr int TK = 64;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows [row0, row0+TQ) of one head's 80-wide slab -> smem tile, zero beyond n_valid
__device__ __forceinline__ void load_tile(__nv_bfloat16 (*dst)[HDP], const __nv_bfloat16* src, int64_t ld, int n_valid) {
  for (int idx = threadIdx.x; idx < TQ * (HD / 8); idx += blockDim.x) {
    const int row = idx / (HD / 8);
    const int ch = idx % (HD / 8);
    if (row < n_valid)
      cp_async16(&dst[row][ch * 8], src + row * ld + ch * 8);
    else
      *reinterpret_cast<uint4*>(&dst[row][ch * 8]) = make_uint4(0, 0, 0, 0);
  }
}

// SPLIT ("precise" mode): q/k/v arrive as hi + lo bf16 pairs and the probabilities are
// split the same way; S = Qh Kh + Qh Kl + Ql Kh and O += Ph Vh + Ph Vl + Pl Vh (the lo x lo terms are
// below 2^-16 relative), so every operand carries ~16 mantissa bits at 3x the MMA work.
using Tile = __nv_bfloat16 (*)[HDP];

template <bool SPLIT>
__global__ void __launch_bounds__(128)
attn_varlen_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ qkv_lo,
                   __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_lo,
                   const int4* __restrict__ tiles, int64_t ld_qkv, int64_t ld_out, int hidden, float scale_log2) {
  extern __shared__ __align__(16) uint8_t attn_smem[];
  constexpr int kTile = TQ * HDP * 2;
  Tile sQ = reinterpret_cast<Tile>(attn_smem);
  Tile sK = reinterpret_cast<Tile>(attn_smem + kTile);
  Tile sV = reinterpret_cast<Tile>(attn_smem + 2 * kTile);
  Tile sQl = reinterpret_cast<Tile>(attn_smem + 3 * kTile);   // SPLIT only
  Tile sKl = reinterpret_cast<Tile>(attn_smem + 4 * kTile);
  Tile sVl = reinterpret_cast<Tile>(attn_smem + 5 * kTile);

  const int4 tile = tiles[blockIdx.x];
  const int q_start = tile.x, q_len = tile.y, kv_start = tile.z, kv_len = tile.w;
  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* qbase = qkv + static_cast<int64_t>(q_start) * ld_qkv + head * HD;
  const __nv_bfloat16* kbase = qkv + static_cast<int64_t>(kv_start) * ld_qkv + hidden + head * HD;
  const __nv_bfloat16* vbase = kbase + hidden;

  const int64_t lo_off = SPLIT ? qkv_lo - qkv : 0;   // same layout, element offset between the two buffers
  load_tile(sQ, qbase, ld_qkv, q_len);
  if constexpr (SPLIT) load_tile(sQl, qbase + lo_off, ld_qkv, q_len);

  uint32_t qf[HD / 16][4];
  uint32_t ql[SPLIT ? HD / 16 : 1][4];
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int kt = 0; kt < kv_len; kt += TK) {
    if (kt > 0) __syncthreads();  // everyone is done with the previous K/V tile
    const int nk = min(TK, kv_len - kt);
    load_tile(sK, kbase + static_cast<int64_t>(kt) * ld_qkv, ld_qkv, nk);
    load_tile(sV, vbase + static_cast<int64_t>(kt) * ld_qkv, ld_qkv, nk);
    if constexpr (SPLIT) {
      load_tile(sKl, kbase + lo_off + static_cast<int64_t>(kt) * ld_qkv, ld_qkv, nk);
      load_tile(sVl, vbase + lo_off + static_cast<int64_t>(kt) * ld_qkv, ld_qkv, nk);
    }
    cp_async_wait_all();
    __syncthreads();
    if (kt == 0) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks)
        ldsm_x4(qf[ks], &sQ[warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
      if constexpr (SPLIT) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks)
          ldsm_x4(ql[ks], &sQl[warp * 16 + (lane & 15)][ks * 16 + (lane >> 4) * 8]);
      }
    }

    // S = Q K^T  (16 x 64 per warp)
    float s[TK / 8][4];
#pragma unroll
    for (int i = 0; i < TK / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int j = 0; j < TK / 16; ++j) {
        uint32_t b[4];
        ldsm_x4(b, &sK[j * 16 + (lane & 7) + (lane >> 4) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
        mma_bf16(s[2 * j], qf[ks], b[0], b[1]);
        mma_bf16(s[2 * j + 1], qf[ks], b[2], b[3]);
        if constexpr (SPLIT) {
          mma_bf16(s[2 * j], ql[ks], b[0], b[1]);
          mma_bf16(s[2 * j + 1], ql[ks], b[2], b[3]);
          ldsm_x4(b, &sKl[j * 16 + (lane & 7) + (lane >> 4) * 8][ks * 16 + ((lane >> 3) & 1) * 8]);
          mma_bf16(s[2 * j], qf[ks], b[0], b[1]);
          mma_bf16(s[2 * j + 1], qf[ks], b[2], b[3]);
        }
      }
    }

    // scale, mask the ragged tail, online softmax (fp32)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < TK / 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = i * 8 + 2 * t + (e & 1);
        const float v = (col < nk) ? s[i][e] * scale_log2 : -INFINITY;
        s[i][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(m_run[h], mx[h]);
      corr[h] = exp2f(m_run[h] - m_new);
      m_run[h] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < TK / 8; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[i][e] - m_run[e >> 1]);
        s[i][e] = pv;
        rs[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 1);
      rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 2);
      l_run[h] = l_run[h] * corr[h] + rs[h];
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }

    // O += P V   (P: 16 x 64 bf16 from the S fragments, V: 64 x 80)
#pragma unroll
    for (int ks = 0; ks < TK / 16; ++ks) {
      uint32_t a[4], al[4];
      if constexpr (SPLIT) {
        split_bf16x2(s[2 * ks][0], s[2 * ks][1], a[0], al[0]);
        split_bf16x2(s[2 * ks][2], s[2 * ks][3], a[1], al[1]);
        split_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1], a[2], al[2]);
        split_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3], a[3], al[3]);
      } else {
        a[0] = pack_bf16x2(s[2 * ks][0], s[2 * ks][1]);
        a[1] = pack_bf16x2(s[2 * ks][2], s[2 * ks][3]);
        a[2] = pack_bf16x2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
        a[3] = pack_bf16x2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
      }
#pragma unroll
      for (int dn = 0; dn < HD / 16; ++dn) {
        uint32_t b[4];
        ldsm_x4_t(b, &sV[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dn * 16 + (lane >> 4) * 8]);
        mma_bf16(o[2 * dn], a, b[0], b[1]);
        mma_bf16(o[2 * dn + 1], a, b[2], b[3]);
        if constexpr (SPLIT) {
          mma_bf16(o[2 * dn], al, b[0], b[1]);
          mma_bf16(o[2 * dn + 1], al, b[2], b[3]);
          ldsm_x4_t(b, &sVl[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dn * 16 + (lane >> 4) * 8]);
          mma_bf16(o[2 * dn], a, b[0], b[1]);
          mma_bf16(o[2 * dn + 1], a, b[2], b[3]);
        }
      }
    }
  }

  // normalise, stage this warp's 16 x 80 block in its own (now idle) Q rows, store 16 B per lane
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) {
    uint32_t h0, h1, l0 = 0, l1 = 0;
    if constexpr (SPLIT) {
      split_bf16x2(o[i][0] * inv0, o[i][1] * inv0, h0, l0);
      split_bf16x2(o[i][2] * inv1, o[i][3] * inv1, h1, l1);
      *reinterpret_cast<uint32_t*>(&sQl[warp * 16 + g][i * 8 + 2 * t]) = l0;
      *reinterpret_cast<uint32_t*>(&sQl[warp * 16 + g + 8][i * 8 + 2 * t]) = l1;
    } else {
      h0 = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
      h1 = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
    }
    *reinterpret_cast<uint32_t*>(&sQ[warp * 16 + g][i * 8 + 2 * t]) = h0;
    *reinterpret_cast<uint32_t*>(&sQ[warp * 16 + g + 8][i * 8 + 2 * t]) = h1;
  }
  __syncwarp();
  for (int idx = lane; idx < 16 * (HD / 8); idx += 32) {
    const int row = warp * 16 + idx / (HD / 8);
    const int ch = idx % (HD / 8);
    if (row < q_len) {
      const int64_t off = static_cast<int64_t>(q_start + row) * ld_out + head * HD + ch * 8;
      *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(&sQ[row][ch * 8]);
      if constexpr (SPLIT) *reinterpret_cast<uint4*>(out_lo + off) = *reinterpret_cast<const uint4*>(&sQl[row][ch * 8]);
    }
  }
}

}  // namespace

cudaError_t attention_launch(const __nv_bfloat16* qkv, const __nv_bfloat16* qkv_lo, __nv_bfloat16* out,
                             __nv_bfloat16* out_lo, const int4* tiles, int n_tiles, int num_heads, int64_t ld_qkv,
                             int64_t ld_out, cudaStream_t stream) {
  if (n_tiles <= 0) return cudaSuccess;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid(n_tiles, num_heads);
  constexpr int kTile = TQ * HDP * 2;
  if (qkv_lo != nullptr) {
    if (out_lo == nullptr) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(attn_varlen_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kTile);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    attn_varlen_kernel<true><<<grid, 128, 6 * kTile, stream>>>(qkv, qkv_lo, out, out_lo, tiles, ld_qkv, ld_out,
                                                               num_heads * HD, scale_log2);
  } else {
    // single-operand attention runs on the tcgen05 kernel (attention2.cu); this mma.sync kernel is kept for the hi + lo
    // operands of the split precision only
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace vr2
