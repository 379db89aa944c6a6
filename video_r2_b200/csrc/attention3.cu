// The split ("hi + lo bf16 operands") precision's attention on the 5th-generation tensor cores: the same
// variable-length, non-causal attention as attention2.cu (HF modeling_qwen2_5_vl.py:221-287), with every operand carried
// as a bf16 pair hi + lo (lo = the rounding residual of hi, ~16 significand bits together):
//
//     S = Qh Kh^T + Qh Kl^T + Ql Kh^T              O += Ph Vh + Ph Vl + Pl Vh
//
// (the lo x lo terms are below 2^-16 relative): three tcgen05.mma batches into ONE TMEM accumulator per contraction.
// Six operand tiles per key block do not fit beside the two-stream layout of attention2.cu, so this kernel has its own
// shape: ONE item (128-row query tile, head) in flight per CTA, key blocks of 64 rows, a 3-stage ring of
// K_hi | K_lo | V_hi | V_lo (40 KB per stage), P_hi | P_lo [128 x 64] in shared memory.
//   warp 4            TMA producer (Q hi + lo per item; K|V hi + lo per block)
//   warp 5            tcgen05.mma issuer: S = 3 x (4 + 1) MMAs of 128x64x16; O += 3 x 4 x {N=64, N=16} MMAs, V MN-major;
//                     S(j+1) is issued right behind P V(j)
//   warps 0..3        softmax, thread = query row = TMEM lane: row max, exp2, P -> (hi, lo) bf16 into the K-major
//                     128B-swizzled layout, running sum; O accumulates in TMEM with the lazily raised reference maximum
//                     of attention2.cu; after the last block O -> registers -> normalised -> (hi, lo) -> global.
// Replaces the mma.sync kernel of round 1 (csrc/attention.cu): no HMMA is left in the library.
#include "attention.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace vr2 {

namespace {

constexpr int HD = 80;
constexpr int TQ = 128;
constexpr int TK = 64;
constexpr int KV_STAGES = 3;
constexpr int Q64 = TQ * 128, Q16 = TQ * 32, QT = Q64 + Q16;        // a [128 x 80] tile: 64-column SW128 box + 16-column SW32 box
constexpr int K64 = TK * 128, K16 = TK * 32, KT = K64 + K16;        // a [64 x 80] tile
constexpr int STAGE = 4 * KT;                                        // K_hi | K_lo | V_hi | V_lo
constexpr int PB = TQ * 128;                                         // [128 x 64 bf16]: one SW128 atom column

constexpr int OFF_Q = 0;                        // Q_hi | Q_lo
constexpr int OFF_KV = OFF_Q + 2 * QT;
constexpr int OFF_P = OFF_KV + KV_STAGES * STAGE;   // P_hi | P_lo
constexpr int OFF_BAR = OFF_P + 2 * PB;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(OFF_KV % 1024 == 0 && KT % 1024 == 0 && OFF_P % 1024 == 0 && QT % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");

constexpr int TM_S = 0;      // S: columns [0, 64)
constexpr int TM_O = 64;     // O: columns [64, 144)
constexpr int TM_COLS = 256;

__device__ __forceinline__ uint64_t desc_k_sw32(uint32_t addr) {      // K-major, 32-byte rows, 8-row groups every 256 B
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr) {    // MN-major, 128-byte rows, 8-row atoms every 1024 B
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t addr) {     // MN-major, 32-byte rows, atoms every 256 B
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kThreads = 192;   // 4 softmax warps + producer + MMA issuer

struct Maps {
  CUtensorMap q64[2], q16[2];   // [hi, lo] of the qkv tensor, 128-row boxes (Q)
  CUtensorMap k64[2], k16[2];   // 64-row boxes (K and V)
};

__global__ void __launch_bounds__(kThreads, 1)
attn_split_kernel(const __grid_constant__ Maps tm, const int2* __restrict__ seg, int n_rows, int num_heads, int hidden,
                  float scale_log2, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int64_t ld_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_empty = q_full + 1;
  uint64_t* kv_full = q_empty + 1;            // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;   // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (n_rows + TQ - 1) / TQ;
  const int n_items = n_tiles * num_heads;

  if (warp == 4 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      tma_prefetch_desc(&tm.q64[i]);
      tma_prefetch_desc(&tm.q16[i]);
      tma_prefetch_desc(&tm.k64[i]);
      tma_prefetch_desc(&tm.k16[i]);
    }
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);   // one arrive per softmax warp
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, TM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // key block range of a query tile: from the segment start of its first row to the segment end of its last
  auto kv_range = [&](int item, int& kv_lo, int& nblk) {
    const int r0 = (item / num_heads) * TQ;
    const int r1 = min(r0 + TQ, n_rows) - 1;
    kv_lo = seg[r0].x;
    nblk = (seg[r1].y - kv_lo + TK - 1) / TK;
  };

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer (whole warp; one elected lane issues)
    const bool issuer = elect_one();
    uint32_t it = 0, n_item = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_item) {
      const int tile = item / num_heads, head = item % num_heads;
      int kv_lo, nblk;
      kv_range(item, kv_lo, nblk);
      mbar_wait(q_empty, (n_item & 1u) ^ 1u);     // the previous item's S MMAs have read Q
      if (issuer) {
        mbar_expect_tx(q_full, 2 * QT);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint8_t* q = smem + OFF_Q + h * QT;
          tma_load_2d(q, &tm.q64[h], q_full, head * HD, tile * TQ);
          tma_load_2d(q + Q64, &tm.q16[h], q_full, head * HD + 64, tile * TQ);
        }
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j, ++it) {
        const uint32_t s = it % KV_STAGES, ph = (it / KV_STAGES) & 1u;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        uint8_t* kb = smem + OFF_KV + s * STAGE;
        const int row = kv_lo + j * TK;
        if (issuer) {
          mbar_expect_tx(&kv_full[s], STAGE);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint8_t* kt = kb + h * KT;              // K_hi, K_lo
            uint8_t* vt = kb + (2 + h) * KT;        // V_hi, V_lo
            tma_load_2d(kt, &tm.k64[h], &kv_full[s], hidden + head * HD, row);
            tma_load_2d(kt + K64, &tm.k16[h], &kv_full[s], hidden + head * HD + 64, row);
            tma_load_2d(vt, &tm.k64[h], &kv_full[s], 2 * hidden + head * HD, row);
            tma_load_2d(vt + K64, &tm.k16[h], &kv_full[s], 2 * hidden + head * HD + 64, row);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer (whole warp; one elected lane issues)
    const bool issuer = elect_one();
    constexpr uint32_t id_s = idesc(TQ, TK, false);
    constexpr uint32_t id_o64 = idesc(TQ, 64, true);
    constexpr uint32_t id_o16 = idesc(TQ, 16, true);
    uint32_t it = 0, n_item = 0, n_pv = 0;
    const uint32_t q_addr = smem_u32(smem + OFF_Q);
    const uint32_t p_addr = smem_u32(smem + OFF_P);
    const uint32_t d_s = tmem_base + TM_S, d_o = tmem_base + TM_O;
    auto issue_s = [&](uint32_t s, bool last) {   // S = Qh Kh^T + Qh Kl^T + Ql Kh^T for the block in stage s
      const uint32_t k_addr = smem_u32(smem + OFF_KV + s * STAGE);
      if (issuer) {
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t qa = q_addr + (prod == 2 ? QT : 0);          // Qh, Qh, Ql
          const uint32_t ka = k_addr + (prod == 1 ? KT : 0);          // Kh, Kl, Kh
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_s, umma_desc_sw128(qa + k * 32), umma_desc_sw128(ka + k * 32), id_s, (prod | k) != 0 ? 1u : 0u);
          umma_bf16(d_s, desc_k_sw32(qa + Q64), desc_k_sw32(ka + K64), id_s, 1u);
        }
        umma_commit(s_full);
        if (last) umma_commit(q_empty);   // every S MMA of the item has been issued: Q is free once they retire
      }
      __syncwarp();
    };
    auto issue_pv = [&](uint32_t s, bool first) {   // O (+)= Ph Vh + Ph Vl + Pl Vh
      const uint32_t v_addr = smem_u32(smem + OFF_KV + s * STAGE + 2 * KT);
      if (issuer) {
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
          const uint32_t pa = p_addr + (prod == 2 ? PB : 0);          // Ph, Ph, Pl
          const uint32_t va = v_addr + (prod == 1 ? KT : 0);          // Vh, Vl, Vh
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            const uint64_t da = umma_desc_sw128(pa + kk * 32);
            const uint32_t acc = (prod | kk) != 0 || !first ? 1u : 0u;
            umma_bf16(d_o, da, desc_mn_sw128(va + kk * 2048), id_o64, acc);
            umma_bf16(d_o + 64, da, desc_mn_sw32(va + K64 + kk * 512), id_o16, acc);
          }
        }
        umma_commit(&kv_empty[s]);   // K (read by S earlier) and V of this block are free once these retire
        umma_commit(o_full);
      }
      __syncwarp();
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_item) {
      int kv_lo, nblk;
      kv_range(item, kv_lo, nblk);
      mbar_wait(q_full, n_item & 1u);
      tc_fence_after();
      {
        const uint32_t s = it % KV_STAGES, ph = (it / KV_STAGES) & 1u;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        issue_s(s, nblk == 1);
      }
      for (int j = 0; j < nblk; ++j, ++it, ++n_pv) {
        const uint32_t s = it % KV_STAGES;
        mbar_wait(p_ready, n_pv & 1u);      // P (hi, lo) is in shared memory, S has been read, O rescaled if needed
        tc_fence_after();
        issue_pv(s, j == 0);
        if (j + 1 < nblk) {
          const uint32_t s1 = (it + 1) % KV_STAGES, ph1 = ((it + 1) / KV_STAGES) & 1u;
          mbar_wait(&kv_full[s1], ph1);
          tc_fence_after();
          issue_s(s1, j + 2 == nblk);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps: thread = query row = TMEM lane
    const int tid = threadIdx.x;   // 0..127
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t ts = lane_addr + TM_S, to = lane_addr + TM_O;
    uint8_t* p_hi = smem + OFF_P + tid * 128;
    uint8_t* p_lo = p_hi + PB;
    uint32_t nblk_done = 0;
    constexpr float kTau = 8.0f;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int tile = item / num_heads, head = item % num_heads;
      int kv_lo, nblk;
      kv_range(item, kv_lo, nblk);
      const int r = tile * TQ + tid;
      int lo = 0, hi = 0;
      if (r < n_rows) {
        const int2 sg = seg[r];
        lo = sg.x;
        hi = sg.y;
      }
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j) {
        const uint32_t b = nblk_done + j;
        const int c_lo = lo - (kv_lo + j * TK), c_hi = hi - (kv_lo + j * TK);   // this row's valid columns of the block
        mbar_wait(s_full, b & 1u);
        tc_fence_after();
        const bool outside = c_hi <= 0 || c_lo >= TK;
        const bool skip = __all_sync(0xffffffffu, outside);      // tcgen05.ld is warp-collective
        const bool interior = c_lo <= 0 && c_hi >= TK;
        uint32_t sv[TK];
        float mx = -INFINITY;
        if (!skip) {
          tmem_ld32(ts, sv);
          tmem_ld32(ts + 32, sv + 32);
          tmem_ld_wait();
          if (outside) {
          } else if (interior) {
#pragma unroll
            for (int i = 0; i < TK; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
          } else {
#pragma unroll
            for (int i = 0; i < TK; ++i)
              if (i >= c_lo && i < c_hi) mx = fmaxf(mx, __uint_as_float(sv[i]));
          }
        }
        if (j == 0) {
          m_ref = mx;
        } else {
          // S_j complete => P V(j-1), issued before it, has retired: O holds blocks 0 .. j-1 and may be rescaled
          mbar_wait(o_full, (b - 1) & 1u);
          const bool need = (mx - m_ref) * scale_log2 > kTau || (m_ref == -INFINITY && mx != -INFINITY);
          if (__any_sync(0xffffffffu, need)) {
            tc_fence_after();
            const float m_new = fmaxf(m_ref, mx);
            const float f = m_ref == -INFINITY ? 0.f : ex2_approx((m_ref - m_new) * scale_log2);
#pragma unroll 1
            for (int c = 0; c < HD; c += 16) {
              uint32_t v[16];
              tmem_ld16(to + c, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st16(to + c, v);
            }
            tmem_st_wait();
            l_run *= f;
            m_ref = m_new;
          }
        }
        const float m_scaled = m_ref == -INFINITY ? 0.f : m_ref * scale_log2;
        // probabilities -> (hi, lo) bf16 -> shared memory, K-major SWIZZLE_128B (8 chunks of 16 bytes per row)
        float rs = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t ph[4], pl[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = q * 8 + 2 * e;
            float p0 = 0.f, p1 = 0.f;
            if (!skip && !outside) {
              if (interior || (i >= c_lo && i < c_hi)) p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), scale_log2, -m_scaled));
              if (interior || (i + 1 >= c_lo && i + 1 < c_hi)) p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), scale_log2, -m_scaled));
            }
            rs += p0 + p1;
            split_bf16x2(p0, p1, ph[e], pl[e]);
          }
          *reinterpret_cast<uint4*>(p_hi + ((q ^ (tid & 7)) << 4)) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
          *reinterpret_cast<uint4*>(p_lo + ((q ^ (tid & 7)) << 4)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
        l_run += rs;
        tc_fence_before();       // this thread's TMEM reads of S (and writes of O) are complete
        fence_proxy_async();     // P visible to the tensor core's shared-memory reads
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
      }
      // the item's last P V has retired: O -> registers -> normalised -> (hi, lo) -> global (160 contiguous bytes per row each)
      mbar_wait(o_full, (nblk_done + nblk - 1) & 1u);
      tc_fence_after();
      nblk_done += nblk;
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll 1
      for (int c = 0; c < HD; c += 16) {
        uint32_t v[16];
        tmem_ld16(to + c, v);
        tmem_ld_wait();
        if (r < n_rows) {
          uint32_t oh[8], ol[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            split_bf16x2(__uint_as_float(v[2 * e]) * inv, __uint_as_float(v[2 * e + 1]) * inv, oh[e], ol[e]);
          const int64_t off = static_cast<int64_t>(r) * ld_out + head * HD + c;
          *reinterpret_cast<uint4*>(out_hi + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
          *reinterpret_cast<uint4*>(out_hi + off + 8) = make_uint4(oh[4], oh[5], oh[6], oh[7]);
          *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
          *reinterpret_cast<uint4*>(out_lo + off + 8) = make_uint4(ol[4], ol[5], ol[6], ol[7]);
        }
      }
      tc_fence_before();   // O may be overwritten by the next item's first P V
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, TM_COLS);
}

}  // namespace

cudaError_t attention_split_launch(const __nv_bfloat16* qkv, const __nv_bfloat16* qkv_lo, __nv_bfloat16* out,
                                   __nv_bfloat16* out_lo, const int2* seg, int n_rows, int num_heads, int64_t ld_qkv,
                                   int64_t ld_out, int num_sms, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  if (qkv_lo == nullptr || out_lo == nullptr || (ld_out % 8) != 0) return cudaErrorInvalidValue;
  const int hidden = num_heads * HD;
  Maps tm;
  const __nv_bfloat16* src[2] = {qkv, qkv_lo};
  for (int h = 0; h < 2; ++h) {
    if (!tma_encode_2d(&tm.q64[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src[h], 3 * hidden, n_rows, ld_qkv * 2, 64, TQ, true) ||
        !tma_encode_2d_swz(&tm.q16[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, src[h], 3 * hidden, n_rows, ld_qkv * 2, 16, TQ, 32) ||
        !tma_encode_2d(&tm.k64[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src[h], 3 * hidden, n_rows, ld_qkv * 2, 64, TK, true) ||
        !tma_encode_2d_swz(&tm.k16[h], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, src[h], 3 * hidden, n_rows, ld_qkv * 2, 16, TK, 32))
      return cudaErrorInvalidValue;
  }
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const int items = ((n_rows + TQ - 1) / TQ) * num_heads;
  const int grid = items < num_sms ? items : num_sms;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  attn_split_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(tm, seg, n_rows, num_heads, hidden, scale_log2, out, out_lo, ld_out);
  return cudaGetLastError();
}

}  // namespace vr2
