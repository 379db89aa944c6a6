// Variable-length, non-causal attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA) for the vision
// tower (HF modeling_qwen2_5_vl.py:221-287): window layers attend inside cu_window_seqlens segments (<= 64 patches),
// layers 7/15/23/31 inside one temporal slice (h*w patches).  q/k/v: bf16, 16 heads x 80, one [N, 3840] buffer
// written by the QKV GEMM epilogue (rotary already applied).
//
// Work item = (128-row query tile, head).  Every patch row r carries the bounds [lo_r, hi_r) of its segment (plan:
// seg_window / seg_full), so one kernel serves both kinds of layer and ragged batches: the tile visits the 128-row
// key/value blocks that cover [lo of its first row, hi of its last row) and each row masks what is outside its own
// segment.  With 64-patch windows a tile is exactly two windows and one key block.
//
// Persistent CTAs, 10 warps, TWO items in flight per CTA (streams 0 and 1, each with its own Q tile, S and O
// accumulators in TMEM, P buffer and softmax warpgroup): while one stream's softmax warps work on their scores the
// tensor core runs the other stream's MMAs, and the loads / stores of one item hide behind the other's compute.
//   warp 8             TMA producer: Q per item, K|V blocks of both streams into one 3-stage ring, in the fixed
//                      (stream, block) order the MMA issuer consumes them.  A 128 x 80 operand tile is two boxes:
//                      64 columns (128-byte swizzle) + 16 columns (32-byte swizzle).
//   warp 9             tcgen05.mma issuer:  S = Q K_j^T (5 MMAs of 128x128x16),  O += P_j V_j (8 k-steps x
//                      {N=64, N=16}, V as an MN-major operand; O accumulates in TMEM over the blocks of an item)
//   warps 4x..4x+3     softmax of stream x, thread = query row = TMEM lane: row max, exp2, P_j -> shared memory
//                      (bf16, the K-major 128B-swizzled layout the MMA reads), running sum.  The exponentials are taken
//                      against a reference maximum that is raised lazily (only when a block's maximum exceeds it by
//                      2^8 in the exponent), and only then is O rescaled — in TMEM, by these warps (tcgen05.ld / st)
//                      between two P V batches.  After the last block O -> registers -> normalised -> shared memory
//                      (aliasing the stream's P buffer) -> one TMA store.
// fp32 scores, softmax and accumulation; bf16 operands (P is rounded to bf16 like q, k, v).
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
constexpr int TK = 128;
constexpr int KV_STAGES = 3;
constexpr int Q64_BYTES = TQ * 128;          // [128 x 64 bf16], SWIZZLE_128B
constexpr int Q16_BYTES = TQ * 32;           // [128 x 16 bf16], SWIZZLE_32B
constexpr int TILE_BYTES = Q64_BYTES + Q16_BYTES;     // 20 KB: one 128 x 80 operand tile
constexpr int KV_STAGE_BYTES = 2 * TILE_BYTES;        // K tile | V tile
constexpr int P_BYTES = 2 * TQ * 128;                 // [128 x 128 bf16] as two 64-column SWIZZLE_128B atoms
// the output staging of a stream (two [128 x 40 bf16] boxes, 80-byte rows = 20 KB) aliases its P buffer: P is dead
// once the item's last P V has retired

constexpr int OFF_Q = 0;                               // [2 streams]
constexpr int OFF_KV = OFF_Q + 2 * TILE_BYTES;
constexpr int OFF_P = OFF_KV + KV_STAGES * KV_STAGE_BYTES;   // [2 streams]
constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB shared-memory limit");

constexpr int TM_S = 0;      // S of stream x: columns [128 x, 128 x + 128)
constexpr int TM_O = 256;    // O of stream x: columns [256 + 128 x, +80)
constexpr int TM_O_STRIDE = 128;

// K-major operand tile with 32-byte rows (16 bf16 = one MMA k-step), SWIZZLE_32B: 8-row groups every 256 bytes
__device__ __forceinline__ uint64_t desc_k_sw32(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;   // SWIZZLE_32B
  return d;
}
// MN-major operand (V as B: rows = key index = MMA K, columns = head dim = MMA N), 128-byte rows of 64 columns,
// SWIZZLE_128B: the 8 k-rows of an atom are 128 bytes apart, atoms along K every 1024 bytes (SBO); a single
// 64-column chunk along N, so LBO is unused.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// same for the 16-column remainder: 32-byte rows, SWIZZLE_32B, atoms along K every 256 bytes
__device__ __forceinline__ uint64_t desc_mn_sw32(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(256 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn_major ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ float ex2_approx(float x) {   // MUFU.EX2: 2 ulp, exact 0 for -inf
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kThreads = 320;   // 2 softmax warpgroups (one per stream) + producer warp + MMA warp

// F16: q, k, v, the probabilities and the output are fp16 instead of bf16 (the f16 activation mode)
template <int F16>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm64, const __grid_constant__ CUtensorMap tm16,
               const __grid_constant__ CUtensorMap tmOut, const int2* __restrict__ seg, int n_rows, int num_heads,
               int hidden, float scale_log2, int reverse) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);   // [2]
  uint64_t* q_empty = q_full + 2;           // [2]
  uint64_t* kv_full = q_empty + 2;          // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES; // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;  // [2]
  uint64_t* p_ready = s_full + 2;           // [2]
  uint64_t* o_full = p_ready + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (n_rows + TQ - 1) / TQ;
  const int n_items = n_tiles * num_heads;
  // query tile of an item: ascending, or (reverse) from the last tile down — see GEMM_REVERSE_M in gemm.cuh
  auto tile_of = [&](int item) { return reverse ? n_tiles - 1 - item / num_heads : item / num_heads; };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm64);
    tma_prefetch_desc(&tm16);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);   // one arrive per softmax warp of the stream
      mbar_init(&o_full[s], 1);
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

  // key/value block range of a query tile: from the segment start of its first row to the segment end of its last
  auto kv_range = [&](int item, int& kv_lo, int& nblk) {
    if (item >= n_items) { kv_lo = 0; nblk = 0; return; }
    const int r0 = tile_of(item) * TQ;
    const int r1 = min(r0 + TQ, n_rows) - 1;
    kv_lo = seg[r0].x;
    nblk = (seg[r1].y - kv_lo + TK - 1) / TK;
  };

  // Items of this CTA are taken in pairs (stream 0, stream 1).  Producer and MMA issuer walk one fixed order of
  // (stream, block): (0,0) (1,0) then, after each P V of (x, j), (x, j+1) — so while one stream's softmax warps work
  // on their scores the tensor core runs the other stream's S / P V, and K|V stages are consumed in the order they
  // were loaded.
  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer (whole warp; one elected lane issues)
    {
      const bool issuer = elect_one();
      uint32_t it = 0, pair_no = 0;
      auto load_q = [&](int x, int item) {
        const int tile = tile_of(item), head = item % num_heads;
        mbar_wait(&q_empty[x], (pair_no & 1u) ^ 1u);     // the stream's previous item's S MMAs have read Q
        uint8_t* q = smem + OFF_Q + x * TILE_BYTES;
        if (issuer) {
          mbar_expect_tx(&q_full[x], TILE_BYTES);
          tma_load_2d(q, &tm64, &q_full[x], head * HD, tile * TQ);
          tma_load_2d(q + Q64_BYTES, &tm16, &q_full[x], head * HD + 64, tile * TQ);
        }
        __syncwarp();
      };
      auto load_kv = [&](int item, int kv_lo, int j) {
        const int head = item % num_heads;
        const uint32_t s = it % KV_STAGES, ph = (it / KV_STAGES) & 1u;
        ++it;
        mbar_wait(&kv_empty[s], ph ^ 1u);
        uint8_t* kb = smem + OFF_KV + s * KV_STAGE_BYTES;
        const int row = kv_lo + j * TK;
        if (issuer) {
          mbar_expect_tx(&kv_full[s], KV_STAGE_BYTES);
          tma_load_2d(kb, &tm64, &kv_full[s], hidden + head * HD, row);
          tma_load_2d(kb + Q64_BYTES, &tm16, &kv_full[s], hidden + head * HD + 64, row);
          tma_load_2d(kb + TILE_BYTES, &tm64, &kv_full[s], 2 * hidden + head * HD, row);
          tma_load_2d(kb + TILE_BYTES + Q64_BYTES, &tm16, &kv_full[s], 2 * hidden + head * HD + 64, row);
        }
        __syncwarp();
      };
      for (int item0 = blockIdx.x; item0 < n_items; item0 += 2 * gridDim.x, ++pair_no) {
        const int item1 = item0 + gridDim.x;
        int lo[2], nb[2];
        kv_range(item0, lo[0], nb[0]);
        kv_range(item1, lo[1], nb[1]);
        const int items[2] = {item0, item1};
        for (int x = 0; x < 2; ++x)
          if (nb[x] > 0) {
            load_q(x, items[x]);
            load_kv(items[x], lo[x], 0);
          }
        const int nmax = max(nb[0], nb[1]);
        for (int j = 0; j < nmax; ++j)
          for (int x = 0; x < 2; ++x)
            if (j + 1 < nb[x]) load_kv(items[x], lo[x], j + 1);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer (whole warp; one elected lane issues)
    {
      const bool issuer = elect_one();
      constexpr uint32_t id_s = umma_idesc_fmt(idesc_bf16(TQ, TK, false), F16 != 0, F16 != 0);
      constexpr uint32_t id_o64 = umma_idesc_fmt(idesc_bf16(TQ, 64, true), F16 != 0, F16 != 0);
      constexpr uint32_t id_o16 = umma_idesc_fmt(idesc_bf16(TQ, 16, true), F16 != 0, F16 != 0);
      uint32_t it = 0;              // K|V ring position of the next block whose S has not been issued
      uint32_t nblk_done[2] = {0, 0};  // blocks of each stream whose S / P V have been issued (barrier phases)
      uint32_t npv_done[2] = {0, 0};
      uint32_t stage_of[2] = {0, 0};   // ring stage holding the stream's current block
      uint32_t pair_no = 0;
      auto issue_s = [&](int x, bool last) {   // S_x = Q_x K^T for the stream's next block
        const uint32_t s = it % KV_STAGES, ph = (it / KV_STAGES) & 1u;
        ++it;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        stage_of[x] = s;
        const uint32_t q_addr = smem_u32(smem + OFF_Q + x * TILE_BYTES);
        const uint32_t k_addr = smem_u32(smem + OFF_KV + s * KV_STAGE_BYTES);
        const uint32_t d = tmem_base + TM_S + x * TK;
        if (issuer) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d, umma_desc_sw128(q_addr + k * 32), umma_desc_sw128(k_addr + k * 32), id_s, k != 0 ? 1u : 0u);
          umma_bf16(d, desc_k_sw32(q_addr + Q64_BYTES), desc_k_sw32(k_addr + Q64_BYTES), id_s, 1u);
          umma_commit(&s_full[x]);
          if (last) umma_commit(&q_empty[x]);   // every S MMA of the item has been issued: Q is free once they retire
        }
        __syncwarp();
        ++nblk_done[x];
      };
      auto issue_pv = [&](int x, bool first) {   // O_x (+)= P_x V for the stream's current block (first: of the item)
        mbar_wait(&p_ready[x], npv_done[x] & 1u);   // P is in shared memory (and O rescaled if the reference maximum moved)
        tc_fence_after();
        ++npv_done[x];
        const uint32_t s = stage_of[x];
        const uint32_t p_addr = smem_u32(smem + OFF_P + x * P_BYTES);
        const uint32_t v_addr = smem_u32(smem + OFF_KV + s * KV_STAGE_BYTES + TILE_BYTES);
        const uint32_t d = tmem_base + TM_O + x * TM_O_STRIDE;
        if (issuer) {
#pragma unroll
          for (int kk = 0; kk < TK / 16; ++kk) {
            const uint64_t da = umma_desc_sw128(p_addr + (kk >> 2) * (TQ * 128) + (kk & 3) * 32);
            umma_bf16(d, da, desc_mn_sw128(v_addr + kk * 2048), id_o64, (kk != 0 || !first) ? 1u : 0u);
            umma_bf16(d + 64, da, desc_mn_sw32(v_addr + Q64_BYTES + kk * 512), id_o16, (kk != 0 || !first) ? 1u : 0u);
          }
          umma_commit(&kv_empty[s]);        // K (read by S earlier) and V of this block are free once these retire
          umma_commit(&o_full[x]);
        }
        __syncwarp();
      };
      for (int item0 = blockIdx.x; item0 < n_items; item0 += 2 * gridDim.x, ++pair_no) {
        const int item1 = item0 + gridDim.x;
        int lo[2], nb[2];
        kv_range(item0, lo[0], nb[0]);
        kv_range(item1, lo[1], nb[1]);
        for (int x = 0; x < 2; ++x)
          if (nb[x] > 0) {
            mbar_wait(&q_full[x], pair_no & 1u);
            issue_s(x, nb[x] == 1);
          }
        const int nmax = max(nb[0], nb[1]);
        for (int j = 0; j < nmax; ++j)
          for (int x = 0; x < 2; ++x)
            if (j < nb[x]) {
              issue_pv(x, j == 0);
              if (j + 1 < nb[x]) issue_s(x, j + 2 == nb[x]);
            }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups: stream x = warp / 4,
    //                                                              thread = query row = TMEM lane
    const int x = warp >> 2;
    const int tid = threadIdx.x & 127;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t ts = lane_addr + TM_S + x * TK;
    const uint32_t to = lane_addr + TM_O + x * TM_O_STRIDE;
    uint8_t* pbuf = smem + OFF_P + x * P_BYTES;
    uint32_t nblk_done = 0;   // blocks of this stream processed so far (barrier phases)
    for (int item = blockIdx.x + x * gridDim.x; item < n_items; item += 2 * gridDim.x) {
      const int tile = tile_of(item), head = item % num_heads;
      int kv_lo, nblk;
      kv_range(item, kv_lo, nblk);
      const int r = tile * TQ + tid;
      int lo = 0, hi = 0;
      if (r < n_rows) {
        const int2 sg = seg[r];
        lo = sg.x;
        hi = sg.y;
      }
      // O accumulates in TMEM across the key blocks of the item (the P V MMAs of block j > 0 accumulate).  The
      // exponentials of a row are taken against m_ref, which follows the running maximum LAZILY: it moves — and O and
      // the running sum are rescaled, O in TMEM — only when a block's maximum exceeds it by more than kTau in the
      // exponent (P then stays <= 2^kTau, far inside bf16 / fp16 range, sums and O are fp32).  After the first block
      // or two the maximum rarely grows that much, so the per-block cost is the softmax itself, not an 80-column
      // read-modify-write of O.
      constexpr float kTau = 8.0f;
      float m_ref = -INFINITY, l_run = 0.f;

      // the previous item's TMA store has read the staging area that aliases this stream's P buffer
      if (tid == 0) tma_store_wait_read<0>();
      named_bar_sync(1 + x, 128);

      for (int j = 0; j < nblk; ++j) {
        const uint32_t b = nblk_done + j;
        const int c_lo = lo - (kv_lo + j * TK), c_hi = hi - (kv_lo + j * TK);   // this row's valid columns of the block
        mbar_wait(&s_full[x], b & 1u);
        tc_fence_after();
        // Each 64-column half of the block is, for this row, entirely inside its segment (no masking), entirely
        // outside (skipped: P = 0) or cut by a segment boundary (masked element by element).
        // pass 1: row maximum over the valid columns (64 columns per TMEM round trip)
        float mx = -INFINITY;
#pragma unroll 1
        for (int c0 = 0; c0 < TK; c0 += 64) {
          // tcgen05.ld is warp-collective: a half is skipped only when it is outside for every row of the warp
          const bool outside = c_hi <= c0 || c_lo >= c0 + 64;
          if (__all_sync(0xffffffffu, outside)) continue;
          const bool interior = c_lo <= c0 && c_hi >= c0 + 64;
          uint32_t v[64];
          tmem_ld32(ts + c0, v);
          tmem_ld32(ts + c0 + 32, v + 32);
          tmem_ld_wait();
          if (outside) {
          } else if (interior) {
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
            for (int i = 4; i < 64; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
            }
            mx = fmaxf(mx, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              const int c = c0 + i;
              if (c >= c_lo && c < c_hi) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
          }
        }
        if (j == 0) {
          m_ref = mx;
        } else {
          // S_j complete => the P V of block j - 1 (issued before it) has retired: O holds blocks 0 .. j-1 and may be
          // rescaled; observing o_full here also keeps this thread in step with the barrier's phases
          mbar_wait(&o_full[x], (b - 1) & 1u);
          const bool need = (mx - m_ref) * scale_log2 > kTau || (m_ref == -INFINITY && mx != -INFINITY);
          if (__any_sync(0xffffffffu, need)) {          // tcgen05.ld / st are warp-collective
            tc_fence_after();
            const float m_new = fmaxf(m_ref, mx);
            const float f = m_ref == -INFINITY ? 0.f : ex2_approx((m_ref - m_new) * scale_log2);
#pragma unroll 1
            for (int c = 0; c < HD; c += 16) {       // rare path
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
        // exp2((s - m_ref) * scale): scores stay unscaled in TMEM, the 1/sqrt(d) * log2(e) factor is applied here
        const float m_scaled = m_ref == -INFINITY ? 0.f : m_ref * scale_log2;
        // pass 2: probabilities -> bf16 -> shared memory in the K-major SWIZZLE_128B layout the P V MMAs read
        uint8_t* prow = pbuf + tid * 128;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < TK; c0 += 32) {
          uint8_t* atom = prow + (c0 >> 6) * (TQ * 128);     // 64 key columns = one swizzle atom row of 128 bytes
          const int ch0 = (c0 & 63) >> 3;
          const bool outside = c_hi <= c0 || c_lo >= c0 + 32;
          if (__all_sync(0xffffffffu, outside)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(atom + (((ch0 + q) ^ (tid & 7)) << 4)) = make_uint4(0, 0, 0, 0);
            continue;
          }
          const bool interior = c_lo <= c0 && c_hi >= c0 + 32;
          uint32_t v[32];
          tmem_ld32(ts + c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (outside) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          } else if (interior) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), scale_log2, -m_scaled));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m_scaled));
              rs0 += p0;
              rs1 += p1;
              pk[i >> 1] = pack16x2<F16>(p0, p1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const int c = c0 + i;
              const float p0 = (c >= c_lo && c < c_hi) ? ex2_approx(fmaf(__uint_as_float(v[i]), scale_log2, -m_scaled)) : 0.f;
              const float p1 = (c + 1 >= c_lo && c + 1 < c_hi) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), scale_log2, -m_scaled)) : 0.f;
              rs0 += p0;
              rs1 += p1;
              pk[i >> 1] = pack16x2<F16>(p0, p1);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(atom + (((ch0 + q) ^ (tid & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        l_run += rs0 + rs1;
        tc_fence_before();       // this thread's TMEM reads of S (and writes of O) are complete
        fence_proxy_async();     // P visible to the tensor core's shared-memory reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[x]);
      }
      // the item's last P V has retired: O (80 fp32 columns of this row) -> registers
      float o[HD];
      {
        mbar_wait(&o_full[x], (nblk_done + nblk - 1) & 1u);
        tc_fence_after();
        uint32_t v[HD];
        tmem_ld32(to, v);
        tmem_ld32(to + 32, v + 32);
        tmem_ld16(to + 64, v + 64);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < HD; ++i) o[i] = __uint_as_float(v[i]);
        tc_fence_before();
      }
      nblk_done += nblk;

      // ---- normalise, stage [128 x 80] bf16 (two 40-column boxes, 80-byte rows: conflict-free) in the P buffer
      //      (its last P V has retired: o_full was just waited on), one TMA store
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      named_bar_sync(1 + x, 128);              // every row of the stream has finished reading O / is past its P writes
      uint8_t* orow = pbuf + tid * 80;
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
        tma_store_2d(&tmOut, pbuf, head * HD, tile * TQ);
        tma_store_2d(&tmOut, pbuf + TQ * 80, head * HD + 40, tile * TQ);
        tma_store_commit();
      }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

}  // namespace

cudaError_t attention_tc_launch(const __nv_bfloat16* qkv, __nv_bfloat16* out, const int2* seg, int n_rows, int num_heads,
                                int64_t ld_qkv, int64_t ld_out, int num_sms, cudaStream_t stream, bool f16, bool reverse) {
  if (n_rows <= 0) return cudaSuccess;
  const int hidden = num_heads * HD;
  CUtensorMap tm64, tm16, tmOut;
  if (!tma_encode_2d(&tm64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3 * hidden, n_rows, ld_qkv * 2, 64, TQ, true) ||
      !tma_encode_2d_swz(&tm16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qkv, 3 * hidden, n_rows, ld_qkv * 2, 16, TQ, 32) ||
      !tma_encode_2d(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, hidden, n_rows, ld_out * 2, 40, TQ, false))
    return cudaErrorInvalidValue;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(attn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const int items = ((n_rows + TQ - 1) / TQ) * num_heads;
  const int grid = items < num_sms ? items : num_sms;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  if (f16) attn_tc_kernel<1><<<grid, kThreads, SMEM_BYTES, stream>>>(tm64, tm16, tmOut, seg, n_rows, num_heads, hidden, scale_log2, reverse ? 1 : 0);
  else attn_tc_kernel<0><<<grid, kThreads, SMEM_BYTES, stream>>>(tm64, tm16, tmOut, seg, n_rows, num_heads, hidden, scale_log2, reverse ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace vr2
