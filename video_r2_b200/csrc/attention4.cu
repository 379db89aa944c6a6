// Pipelined variable-length attention on tcgen05 / TMEM / TMA for the long segments of the vision tower: the
// full-attention layers 7/15/23/31 attend inside one temporal slice (h*w patches, HF modeling_qwen2_5_vl.py:221-287,
// :488-502).  Same contract as attention2.cu (q|k|v in one [N, 3*hidden] buffer, rotary applied, per-row segment
// bounds), a different schedule.
//
// What bound the tile kernel (attention2.cu) in these layers (profiles/r02f_attention_summary.md): the per-stream chain
//   softmax(j) -> P V(j) -> S(j+1) -> softmax(j+1)
// — the softmax warps waited for their next score tile a quarter of the time and the MUFU pipe sat at 40 %.  Here
//   * key blocks are 64 rows and every stream has TWO score buffers in TMEM and TWO probability buffers in shared
//     memory: S(j+2) is issued right after P V(j), two blocks ahead of the softmax, so a score tile is always waiting
//     when the softmax warps come back, across item boundaries too (S(0), S(1) of the next item are issued while the
//     current item's output is normalised and stored);
//   * the two streams of a CTA are the two adjacent 128-row query tiles of ONE head: they read the same K|V blocks,
//     which are loaded once (half the L2 -> shared-memory traffic) into a 6-stage ring that both streams release;
//   * each stream has its own MMA-issuing warp, so neither waits for the other's softmax (the tile kernel's single
//     issuer walks a fixed (stream, block) order).
// Work item = (pair of adjacent query tiles, head).  The key range of an item is the union of its two tiles' ranges;
// a stream takes part only in the blocks that intersect its own range (so ragged batches and short segments stay
// correct, if not optimal — window layers keep the tile kernel), every row masks what lies outside its own segment.
//
// Warps (11): 0-3 softmax of stream 0, 4-7 softmax of stream 1 (thread = query row = TMEM lane), 8 TMA producer,
// 9 / 10 MMA issuer of stream 0 / 1.  O accumulates in TMEM over the blocks of an item against a lazily raised
// reference maximum, as in attention2.cu.
#include "attention.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gemm.cuh"
#include "ptx.cuh"

namespace vr2 {

namespace {

constexpr int HD = 80;
constexpr int TQ = 128;
constexpr int TK = 64;
constexpr int NS = 6;                              // K|V ring stages
constexpr int Q64_BYTES = TQ * 128;                // [128 x 64] 16-bit, SWIZZLE_128B
constexpr int Q16_BYTES = TQ * 32;                 // [128 x 16] 16-bit, SWIZZLE_32B
constexpr int QTILE = Q64_BYTES + Q16_BYTES;       // 20 KB
constexpr int K64_BYTES = TK * 128;                // [64 x 64]
constexpr int K16_BYTES = TK * 32;                 // [64 x 16]
constexpr int KTILE = K64_BYTES + K16_BYTES;       // 10 KB
constexpr int KV_STAGE = 2 * KTILE;                // K block | V block
constexpr int P_BYTES = TQ * 128;                  // [128 rows x 64 keys] 16-bit, one SWIZZLE_128B atom column

constexpr int OFF_Q = 0;                           // [2 streams]
constexpr int OFF_KV = OFF_Q + 2 * QTILE;
constexpr int OFF_P = OFF_KV + NS * KV_STAGE;      // [2 streams][2 buffers]; a stream's pair also stages its output
constexpr int OFF_BAR = OFF_P + 4 * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 512;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");
static_assert(OFF_KV % 1024 == 0 && KV_STAGE % 1024 == 0 && KTILE % 1024 == 0 && OFF_P % 1024 == 0, "swizzle atoms");

constexpr int TM_S = 0;      // S of stream x, buffer b: columns [128 x + 64 b, +64)
constexpr int TM_O = 256;    // O of stream x: columns [256 + 128 x, +80)

constexpr int kThreads = 352;

// K-major operand tile with 32-byte rows (16 elements = one MMA k-step), SWIZZLE_32B: 8-row groups every 256 bytes
__device__ __forceinline__ uint64_t desc_k_sw32(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
// MN-major operand (V as B: rows = key = MMA K, columns = head dim = MMA N), 128-byte rows, SWIZZLE_128B
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// same for the 16-column remainder: 32-byte rows, SWIZZLE_32B
__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// key blocks of one work item: block j covers rows [lo0 + 64 j, +64); stream x takes part in blocks [jb[x], je[x])
struct Item {
  int lo0, nb;
  int je0, jb1, je1;   // stream 0 starts at block 0
  __device__ __forceinline__ int jb(int x) const { return x ? jb1 : 0; }
  __device__ __forceinline__ int je(int x) const { return x ? je1 : je0; }
};

template <int F16>
__global__ void __launch_bounds__(kThreads, 1)
attn_pipe_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmQ16,
                 const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
                 const __grid_constant__ CUtensorMap tmOut, const int2* __restrict__ seg, int n_rows, int num_heads,
                 int hidden, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [2]
  uint64_t* q_empty = q_full + 2;          // [2]
  uint64_t* kv_full = q_empty + 2;         // [NS]
  uint64_t* kv_empty = kv_full + NS;       // [NS]   two arrivals: one per stream
  uint64_t* s_full = kv_empty + NS;        // [2 streams][2 buffers]
  uint64_t* p_ready = s_full + 4;          // [2][2]  one arrive per softmax warp of the stream
  uint64_t* pv_done = p_ready + 4;         // [2][2]  the P V that read this probability buffer has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (n_rows + TQ - 1) / TQ;
  const int n_pairs = ((n_tiles + 1) / 2) * num_heads;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmQ16);
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmK16);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto item_info = [&](int p, Item& I) {
    I.lo0 = 0; I.nb = 0;
    I.je0 = I.jb1 = I.je1 = 0;
    if (p >= n_pairs) return;
    const int t0 = 2 * (p / num_heads);
    const int r0 = t0 * TQ;
    const int r1 = min(r0 + TQ, n_rows) - 1;
    const int lo0 = seg[r0].x, hi0 = seg[r1].y;
    int hi = hi0;
    I.lo0 = lo0;
    I.je0 = (hi0 - lo0 + TK - 1) / TK;
    if (t0 + 1 < n_tiles) {
      const int r2 = r0 + TQ;
      const int r3 = min(r2 + TQ, n_rows) - 1;
      const int lo1 = seg[r2].x, hi1 = seg[r3].y;
      hi = max(hi0, hi1);
      I.jb1 = (lo1 - lo0) / TK;
      I.je1 = (hi1 - lo0 + TK - 1) / TK;
    }
    I.nb = (hi - lo0 + TK - 1) / TK;
  };

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer (whole warp; one elected lane issues)
    const bool issuer = elect_one();
    uint32_t it = 0, n_q[2] = {0, 0};
    Item cur;
    item_info(blockIdx.x, cur);
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      Item nxt;
      item_info(p + gridDim.x, nxt);
      const int t0 = 2 * (p / num_heads), head = p % num_heads;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (cur.je(x) > cur.jb(x)) {
          mbar_wait(&q_empty[x], (n_q[x] & 1u) ^ 1u);   // the stream's previous item's S MMAs have read Q
          ++n_q[x];
          uint8_t* q = smem + OFF_Q + x * QTILE;
          if (issuer) {
            mbar_expect_tx(&q_full[x], QTILE);
            tma_load_2d(q, &tmQ64, &q_full[x], head * HD, (t0 + x) * TQ);
            tma_load_2d(q + Q64_BYTES, &tmQ16, &q_full[x], head * HD + 64, (t0 + x) * TQ);
          }
          __syncwarp();
        }
      }
      for (int j = 0; j < cur.nb; ++j, ++it) {
        const uint32_t s = it % NS, ph = (it / NS) & 1u;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        uint8_t* kb = smem + OFF_KV + s * KV_STAGE;
        const int row = cur.lo0 + j * TK;
        if (issuer) {
          mbar_expect_tx(&kv_full[s], KV_STAGE);
          tma_load_2d(kb, &tmK64, &kv_full[s], hidden + head * HD, row);
          tma_load_2d(kb + K64_BYTES, &tmK16, &kv_full[s], hidden + head * HD + 64, row);
          tma_load_2d(kb + KTILE, &tmK64, &kv_full[s], 2 * hidden + head * HD, row);
          tma_load_2d(kb + KTILE + K64_BYTES, &tmK16, &kv_full[s], 2 * hidden + head * HD + 64, row);
        }
        __syncwarp();
      }
      cur = nxt;
    }
  } else if (warp == 9 || warp == 10) {
    // ------------------------------------------------------------ MMA issuer of stream x (whole warp; one elected lane)
    const int x = warp - 9;
    const bool issuer = elect_one();
    constexpr uint32_t id_s = umma_idesc_fmt(idesc16(TQ, TK, false), F16 != 0, F16 != 0);
    constexpr uint32_t id_o64 = umma_idesc_fmt(idesc16(TQ, 64, true), F16 != 0, F16 != 0);
    constexpr uint32_t id_o16 = umma_idesc_fmt(idesc16(TQ, 16, true), F16 != 0, F16 != 0);
    const uint32_t q_addr = smem_u32(smem + OFF_Q + x * QTILE);
    const uint32_t kv_addr = smem_u32(smem + OFF_KV);
    const uint32_t p_addr0 = smem_u32(smem + OFF_P + x * 2 * P_BYTES);
    const uint32_t t_s = tmem_base + TM_S + x * 128;
    const uint32_t t_o = tmem_base + TM_O + x * 128;
    uint32_t it0 = 0, n_q = 0, g0 = 0;   // g0: blocks this stream has processed before the current item
    Item cur;
    item_info(blockIdx.x, cur);
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      Item nxt;
      item_info(p + gridDim.x, nxt);
      const int jb = cur.jb(x), je = cur.je(x), n = je - jb;
      auto stage_wait = [&](int j) -> uint32_t {   // block j of this item has landed; returns its ring stage
        const uint32_t g = it0 + static_cast<uint32_t>(j);
        const uint32_t s = g % NS;
        mbar_wait(&kv_full[s], (g / NS) & 1u);
        return s;
      };
      auto skip_block = [&](int j) {   // a block this stream takes no part in: release its stage once it is loaded
        const uint32_t s = stage_wait(j);
        if (issuer) mbar_arrive(&kv_empty[s]);
        __syncwarp();
      };
      for (int j = 0; j < jb; ++j) skip_block(j);
      if (n > 0) {
        // buffers alternate with the stream's running block count g = g0 + t (not with t: items may have odd lengths)
        auto issue_s = [&](int t) {   // S(t) = Q K_t^T into score buffer g & 1
          const uint32_t s = stage_wait(jb + t);
          tc_fence_after();
          const uint32_t k_addr = kv_addr + s * KV_STAGE;
          const uint32_t bs = (g0 + static_cast<uint32_t>(t)) & 1u;
          const uint32_t d = t_s + bs * TK;
          if (issuer) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), id_s, k != 0 ? 1u : 0u);
            umma_bf16(d, desc_k_sw32(q_addr + Q64_BYTES), desc_k_sw32(k_addr + K64_BYTES), id_s, 1u);
            umma_commit(&s_full[x * 2 + bs]);
            if (t == n - 1) umma_commit(&q_empty[x]);   // the item's last S: Q is free once it retires
          }
          __syncwarp();
        };
        mbar_wait(&q_full[x], n_q & 1u);
        ++n_q;
        issue_s(0);
        if (n > 1) issue_s(1);
        for (int t = 0; t < n; ++t) {
          const uint32_t g = g0 + static_cast<uint32_t>(t), b = g & 1u;
          mbar_wait(&p_ready[x * 2 + b], (g >> 1) & 1u);   // P(t) is in shared memory, S(t) has been read, O rescaled if needed
          tc_fence_after();
          const uint32_t s = (it0 + static_cast<uint32_t>(jb + t)) % NS;
          const uint32_t p_addr = p_addr0 + b * P_BYTES;
          const uint32_t v_addr = kv_addr + s * KV_STAGE + KTILE;
          if (issuer) {
#pragma unroll
            for (int kk = 0; kk < TK / 16; ++kk) {
              const uint64_t da = umma_desc_sw128(p_addr + kk * 32);
              umma_bf16(t_o, da, desc_mn_sw128(v_addr + kk * 2048), id_o64, (kk != 0 || t != 0) ? 1u : 0u);
              umma_bf16(t_o + 64, da, desc_mn_sw32(v_addr + K64_BYTES + kk * 512), id_o16, (kk != 0 || t != 0) ? 1u : 0u);
            }
            umma_commit(&kv_empty[s]);          // K (read by S(t) earlier) and V of this block: this stream is done with them
            umma_commit(&pv_done[x * 2 + b]);
          }
          __syncwarp();
          if (t + 2 < n) issue_s(t + 2);        // two blocks ahead of the softmax; its score buffer was freed by p_ready(t)
        }
      }
      for (int j = (n > 0 ? je : 0); j < cur.nb; ++j) skip_block(j);
      it0 += static_cast<uint32_t>(cur.nb);
      g0 += static_cast<uint32_t>(n);
      cur = nxt;
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups: stream x = warp / 4,
    //                                                              thread = query row = TMEM lane
    const int x = warp >> 2;
    const int tid = threadIdx.x & 127;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t t_s = lane_addr + TM_S + x * 128;
    const uint32_t t_o = lane_addr + TM_O + x * 128;
    uint8_t* pbase = smem + OFF_P + x * 2 * P_BYTES;
    uint32_t g0 = 0;   // blocks this stream has processed before the current item; block g uses buffers g & 1
    constexpr float kTau = 8.0f;
    Item cur;
    item_info(blockIdx.x, cur);
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      Item nxt;
      item_info(p + gridDim.x, nxt);
      const int jb = cur.jb(x), n = cur.je(x) - jb;
      if (n > 0) {
        const int tile = 2 * (p / num_heads) + x, head = p % num_heads;
        const int r = tile * TQ + tid;
        int lo = 0, hi = 0;
        if (r < n_rows) {
          const int2 sg = seg[r];
          lo = sg.x;
          hi = sg.y;
        }
        float m_ref = -INFINITY, l_run = 0.f;
        // the previous item's TMA store has read the staging area that aliases this stream's P buffers
        if (tid == 0) tma_store_wait_read<0>();
        named_bar_sync(1 + x, 128);

        // O in TMEM is rescaled (rare) when the reference maximum has to move: P V(g-1), the last one issued, must have
        // retired — O then holds blocks 0 .. t-1 and nothing writes it.  tcgen05.ld / st are warp-collective.
        auto raise_reference = [&](uint32_t g, float mx) {
          mbar_wait(&pv_done[x * 2 + ((g - 1u) & 1u)], ((g - 1u) >> 1) & 1u);
          tc_fence_after();
          const float m_new = fmaxf(m_ref, mx);
          const float f = m_ref == -INFINITY ? 0.f : ex2_approx((m_ref - m_new) * scale_log2);
#pragma unroll 1
          for (int c = 0; c < HD; c += 16) {
            uint32_t o[16];
            tmem_ld16(t_o + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st16(t_o + c, o);
          }
          tmem_st_wait();
          l_run *= f;
          m_ref = m_new;
        };
        bool s_ok = false;   // the next block's score tile was already seen complete (tested early, a block ahead)
        for (int t = 0; t < n; ++t) {
          const uint32_t g = g0 + static_cast<uint32_t>(t), b = g & 1u, ph = (g >> 1) & 1u;
          const int blk0 = cur.lo0 + (jb + t) * TK;
          const int c_lo = lo - blk0, c_hi = hi - blk0;   // this row's valid columns of the block
          if (!s_ok) mbar_wait(&s_full[x * 2 + b], ph);
          tc_fence_after();
          uint8_t* prow = pbase + b * P_BYTES + tid * 128;
          if (__all_sync(0xffffffffu, c_lo <= 0 && c_hi >= TK)) {
            // ---- the block lies inside the segment of every row of the warp (all blocks of an aligned slice):
            // straight-line code.  The exponentials are taken SPECULATIVELY against the current reference maximum while
            // the block maximum is formed beside them (they share no pipe); only if the maximum turns out to exceed the
            // reference by more than kTau (rare after the first block) is the block redone.  Barrier tests are issued
            // early and consumed late so their round trips hide behind the exponentials.
            uint32_t v[TK];
            tmem_ld32(t_s + b * TK, v);
            tmem_ld32(t_s + b * TK + 32, v + 32);
            tmem_ld_wait();
            if (t == 0) {
              float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
              for (int i = 4; i < TK; i += 4) {
                m0 = fmaxf(m0, __uint_as_float(v[i]));
                m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
                m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
              }
              m_ref = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            }
            // the P V that read this probability buffer two blocks ago (or in the previous item) has retired?
            const bool pv_ok = g < 2u || mbar_try_wait(&pv_done[x * 2 + b], ph ^ 1u);
            float rs0, rs1, mx;
            bool p_free = pv_ok;
            auto exps = [&]() {   // probabilities of the block against m_ref -> shared memory; row sum, block maximum
              const float m_scaled = m_ref * scale_log2;   // -inf reference (no valid key so far): redone below
              rs0 = 0.f;
              rs1 = 0.f;
              float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float s0 = __uint_as_float(v[32 * h + i]), s1 = __uint_as_float(v[32 * h + i + 1]);
                  const float p0 = ex2_approx(fmaf(s0, scale_log2, -m_scaled));
                  const float p1 = ex2_approx(fmaf(s1, scale_log2, -m_scaled));
                  m0 = fmaxf(m0, s0);
                  m1 = fmaxf(m1, s1);
                  rs0 += p0;
                  rs1 += p1;
                  pk[i >> 1] = pack16x2<F16>(p0, p1);
                }
                if (!p_free) {
                  mbar_wait(&pv_done[x * 2 + b], ph ^ 1u);
                  p_free = true;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *reinterpret_cast<uint4*>(prow + (((4 * h + q) ^ (tid & 7)) << 4)) =
                      make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
              }
              mx = fmaxf(m0, m1);
            };
            exps();
            if (t > 0) {
              const bool need = (mx - m_ref) * scale_log2 > kTau || m_ref == -INFINITY;
              if (__any_sync(0xffffffffu, need)) {
                raise_reference(g, mx);
                exps();   // overwrites the speculative probabilities (nothing has been published yet)
              }
            }
            // early test of the next block's score tile (issued two blocks ahead of us by the MMA warp)
            s_ok = t + 1 < n && mbar_try_wait(&s_full[x * 2 + (b ^ 1u)], ((g + 1u) >> 1) & 1u);
            l_run += rs0 + rs1;
          } else {
            // ---- a segment boundary cuts the block for some row of the warp (ragged batches, the ends of a slice)
            s_ok = false;
            // tcgen05.ld is warp-collective: the block is skipped only when it is outside for every row of the warp
            const bool skip = __all_sync(0xffffffffu, c_hi <= 0 || c_lo >= TK);
            uint32_t v[TK];
            float mx = -INFINITY;
            if (!skip) {
              tmem_ld32(t_s + b * TK, v);
              tmem_ld32(t_s + b * TK + 32, v + 32);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < TK; ++i)
                if (i >= c_lo && i < c_hi) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            if (t == 0) {
              m_ref = mx;
            } else {
              const bool need = (mx - m_ref) * scale_log2 > kTau || (m_ref == -INFINITY && mx != -INFINITY);
              if (__any_sync(0xffffffffu, need)) raise_reference(g, mx);
            }
            const float m_scaled = m_ref == -INFINITY ? 0.f : m_ref * scale_log2;
            if (g >= 2u) mbar_wait(&pv_done[x * 2 + b], ph ^ 1u);
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const int c0 = 32 * h;
              uint32_t pk[16];
              if (skip) {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const int c = c0 + i;
                  const float s0 = __uint_as_float(h == 0 ? v[i] : v[32 + i]), s1 = __uint_as_float(h == 0 ? v[i + 1] : v[33 + i]);
                  const float p0 = (c >= c_lo && c < c_hi) ? ex2_approx(fmaf(s0, scale_log2, -m_scaled)) : 0.f;
                  const float p1 = (c + 1 >= c_lo && c + 1 < c_hi) ? ex2_approx(fmaf(s1, scale_log2, -m_scaled)) : 0.f;
                  rs0 += p0;
                  rs1 += p1;
                  pk[i >> 1] = pack16x2<F16>(p0, p1);
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(prow + (((4 * h + q) ^ (tid & 7)) << 4)) =
                    make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            l_run += rs0 + rs1;
          }
          tc_fence_before();       // this thread's TMEM reads of S (and writes of O) are complete
          fence_proxy_async();     // P visible to the tensor core's shared-memory reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[x * 2 + b]);
        }

        // the item's last P V has retired (and with it every earlier one): O -> registers
        float o[HD];
        {
          const uint32_t gl = g0 + static_cast<uint32_t>(n) - 1u;
          mbar_wait(&pv_done[x * 2 + (gl & 1u)], (gl >> 1) & 1u);
          tc_fence_after();
          uint32_t w[HD];
          tmem_ld32(t_o, w);
          tmem_ld32(t_o + 32, w + 32);
          tmem_ld16(t_o + 64, w + 64);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < HD; ++i) o[i] = __uint_as_float(w[i]);
          tc_fence_before();
        }
        // normalise, stage [128 x 80] 16-bit (two 40-column boxes, 80-byte rows: conflict-free) over the stream's P
        // buffers (their P V products have retired), one TMA store
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        uint8_t* orow = pbase + tid * 80;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const int i = half * 40 + q * 8;
            *reinterpret_cast<uint4*>(orow + half * (TQ * 80) + q * 16) =
                make_uint4(pack16x2<F16>(o[i] * inv, o[i + 1] * inv), pack16x2<F16>(o[i + 2] * inv, o[i + 3] * inv),
                           pack16x2<F16>(o[i + 4] * inv, o[i + 5] * inv), pack16x2<F16>(o[i + 6] * inv, o[i + 7] * inv));
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + x, 128);
        if (tid == 0) {
          tma_store_2d(&tmOut, pbase, head * HD, tile * TQ);
          tma_store_2d(&tmOut, pbase + TQ * 80, head * HD + 40, tile * TQ);
          tma_store_commit();
        }
        g0 += static_cast<uint32_t>(n);
      }
      cur = nxt;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

AttnKind attention_kind(bool long_segments) {
  const char* e = getenv("VR2_ATTN_KERNEL");
  if (e != nullptr && strcmp(e, "tile") == 0) return ATTN_TILE;
  if (e != nullptr && strcmp(e, "pipe") == 0) return ATTN_PIPE;
  return long_segments ? ATTN_PIPE : ATTN_TILE;
}

cudaError_t attention_launch(AttnKind kind, const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* seg, int n_rows,
                             int num_heads, int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16) {
  switch (kind) {
    case ATTN_PIPE: return attention_pipe_launch(qkv, out, seg, n_rows, num_heads, ld_qkv, ld_out, num_sms, stream, f16);
    default: return attention_tc_launch(qkv, out, seg, n_rows, num_heads, ld_qkv, ld_out, num_sms, stream, f16);
  }
}

cudaError_t attention_pipe_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* seg, int n_rows, int num_heads,
                                  int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16) {
  if (n_rows <= 0) return cudaSuccess;
  const int hidden = num_heads * HD;
  CUtensorMap tmQ64, tmQ16, tmK64, tmK16, tmOut;
  if (!tma_encode_2d(&tmQ64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3 * hidden, n_rows, ld_qkv * 2, 64, TQ, true) ||
      !tma_encode_2d_swz(&tmQ16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * hidden, n_rows, ld_qkv * 2, 16, TQ, 32) ||
      !tma_encode_2d(&tmK64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3 * hidden, n_rows, ld_qkv * 2, 64, TK, true) ||
      !tma_encode_2d_swz(&tmK16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * hidden, n_rows, ld_qkv * 2, 16, TK, 32) ||
      !tma_encode_2d(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, hidden, n_rows, ld_out * 2, 40, TQ, false))
    return cudaErrorInvalidValue;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_pipe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(attn_pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const int n_tiles = (n_rows + TQ - 1) / TQ;
  const int pairs = ((n_tiles + 1) / 2) * num_heads;
  const int grid = pairs < num_sms ? pairs : num_sms;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  if (f16)
    attn_pipe_kernel<1><<<grid, kThreads, SMEM_BYTES, stream>>>(tmQ64, tmQ16, tmK64, tmK16, tmOut, seg, n_rows, num_heads, hidden, scale_log2);
  else
    attn_pipe_kernel<0><<<grid, kThreads, SMEM_BYTES, stream>>>(tmQ64, tmQ16, tmK64, tmK16, tmOut, seg, n_rows, num_heads, hidden, scale_log2);
  return cudaGetLastError();
}

}  // namespace vr2
