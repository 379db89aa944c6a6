// CTA-pair (tcgen05 cta_group::2) GEMM for the three GEMM shapes that carry 97 % of the tower's FLOPs:
//
//   EPI_QKV     C = xb * Wqkv^T      epilogue: * rstd(x), + bias, 2-D rotary on q,k, bf16           (N = 3840, BN = 240)
//   EPI_RESID   C = A  * W^T         epilogue: x += C + bias (fp32), xb = bf16(x * w_norm), sum x^2  (N = 1280, BN = 256)
//   EPI_SWIGLU  C = xb * Wgu^T       epilogue: * rstd, + bias, silu(gate) * up, bf16                 (N = 6912, BN = 256)
//   EPI_LOGPROB C = h  * Wlm^T       epilogue: per-block max / sum-exp of C / T and the target logit: the LM head's
//                                    per-token log-probabilities without materialising [M, vocab] logits
//
// Why a second kernel (ncu r01c, profiles/): in the single-CTA kernel (gemm.cu) the tensor core's operand reads
// take 75 % of the SM's L1/shared data pipe at full rate (128 x 256 tile: 48 KB per 512 cycles) and the epilogue's
// LSU traffic (staging through shared memory, row-strided global accesses) takes the rest: QKV and gate|up sat at
// 100 % of that pipe with the tensor pipe at 51 % / 77 %.  Here
//   * two CTAs of a cluster share one 256 x BN tile: each loads its own 128 rows of A and HALF of W's rows, the
//     leader issues tcgen05.mma.cta_group::2 — operand reads per SM drop to 32 KB per 512 cycles (50 %);
//   * the stage ring shrinks by a third, which pays for epilogue buffers in shared memory;
//   * every epilogue input/output moves by TMA: the fp32 residual tile and the cos/sin rows are bulk-loaded while the
//     MMAs of the tile run, results are written to shared memory once (thread = row, 16-byte stores, layouts chosen
//     conflict-free: 80-byte rows or 128-byte swizzle) and leave through bulk tensor stores issued by a dedicated
//     warp, so epilogue warps issue no global loads/stores except the L1-resident bias / norm-weight vectors.
//
// Warp roles per CTA: E epilogue warps (4 TMEM lane quarters x column groups), 1 TMA producer, 1 MMA issuer (leader
// CTA only), 1 store warp (TMA stores, residual / rotary-table loads, sum-of-squares reduction).
// Replaces the cuBLAS calls behind HF modeling_qwen2_5_vl.py:221-287 (qkv, proj) and :87-88 (gate/up/down).
#include "gemm.cuh"

#include <cstdlib>

#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace vr2 {

namespace {

constexpr int BM = 128;          // accumulator rows per CTA (256 per pair)
constexpr int BK = 64;           // one 128-byte swizzle atom of bf16
constexpr int UK = 16;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kMaxPos = 64;      // positions (patch rows / columns of one frame) the QKV kernel's rotary table holds
constexpr int kBox = 128 * 80;   // one [128 rows x 80 bytes] epilogue box (20 fp32 or 40 bf16 columns)

// V: format of the 16-bit output (and of the activations the epilogue produces): 0 bf16, 1 fp16 (GEMM_OUT_F16).
template <int EPI, int V>
struct Cfg2;

template <int V>
struct Cfg2<EPI_QKV, V> {
  static constexpr int BN = 240, groups = 3, stages = 5;
  static constexpr int out_bytes = 6 * kBox;   // 3 heads x (first | second rotary half) boxes of 40 bf16 columns
  static constexpr int aux_bytes = kMaxPos * 160;   // cos | sin of every (h or w) position: [64][40] fp32
};
template <int V>
struct Cfg2<EPI_LOGPROB, V> {
  static constexpr int BN = 256, groups = 4, stages = 6;
  static constexpr int out_bytes = 0;
  static constexpr int aux_bytes = 2 * 4 * 128 * 8;   // two item parities x 4 column groups x 128 rows x (max, sum)
};
template <int V>
struct Cfg2<EPI_SWIGLU, V> {
  static constexpr int BN = 256, groups = 4, stages = 6;
  static constexpr int out_bytes = 2 * 16384;  // 128 x 128 bf16 as two 128B-swizzled 64-column boxes (single buffer: its
                                               // store drains long before the next tile's accumulator is complete)
  static constexpr int aux_bytes = 0;
};
template <int V>
struct Cfg2<EPI_MROPE, V> {
  static constexpr int BN = 256, groups = 4, stages = 5;   // two heads of 128 per tile
  static constexpr int out_bytes = 4 * 16384;              // 128 x 256 bf16 as four 128B-swizzled 64-column boxes
  static constexpr int aux_bytes = 0;
};
// EPI_RESID moves its 128 x 256 tile through shared memory in sixteen sub-steps of 16 columns over four 12 KB buffers
// (fp32 residual: one [128 x 16] box with 64-byte rows, SWIZZLE_64B, updated in place; 16-bit operand: one [128 x 16]
// box with 32-byte rows, SWIZZLE_32B), each buffer served by its own group of 4 epilogue warps (one per TMEM lane
// quarter): while one buffer is computed on, the others' stores drain and their next residual pieces load.
// Round 2: the sub-steps were halved (32 -> 16 columns) to pay for a FIFTH ring stage — the residual GEMMs stream the
// fp32 tile in and out beside the operands, the loads take longer to land, and with four stages the MMA warp ran dry
// (3 stages: down +18 %, proj +8 %; profiles/r02l_summary.md).
// History (profiles/): whole-tile buffers with 128 x 160 tiles left a 4-stage ring of short stages and a serial
// load -> compute -> store chain (tensor pipe 29 % / 43 %); 256-wide tiles keep the operand traffic per FLOP of the
// gate|up GEMM, small buffers keep a 4-stage ring of 32 KB stages.
// Two shapes of the same epilogue (internal id EPI_RESID_DEEP = EPI_RESID with the deep ring):
//   EPI_RESID       32-column sub-steps (8 per tile; 128-byte fp32 rows / 64-byte 16-bit rows), 4-stage ring — for short
//                   K (proj, K = 1280: 20 k-blocks per tile), where the epilogue's residual traffic is the long pole and
//                   wide TMA boxes serve it best;
//   EPI_RESID_DEEP  16-column sub-steps (16 per tile; 64- / 32-byte rows), 5-stage ring — for long K (down, K = 3456),
//                   where the MMA warp running dry on a 4-stage ring was the long pole.
// A/B in one session (profiles/r02l_summary.md): proj 80.3 us with 32-column sub-steps vs 90.6 with 16; down 149.7 vs 136.1.
constexpr int EPI_RESID_DEEP = 100;
__host__ __device__ constexpr int sub_cols(int epi) { return epi == EPI_RESID_DEEP ? 16 : 32; }      // columns per sub-step
__host__ __device__ constexpr int sub_steps(int epi) { return 256 / sub_cols(epi); }                  // sub-steps per 256-column tile
__host__ __device__ constexpr int sub_x_bytes(int epi) { return 128 * sub_cols(epi) * 4; }            // [128 rows x cols fp32], swizzle span = row bytes
__host__ __device__ constexpr int sub_bytes(int epi) { return sub_x_bytes(epi) + 128 * sub_cols(epi) * 2; }   // + [128 rows x cols 16-bit]
constexpr int kResidBufs = 4;
template <int EPI> constexpr bool is_resid = EPI == EPI_RESID || EPI == EPI_RESID_DEEP;
// 16-byte chunk swizzle of a row of RB bytes as TMA SWIZZLE_{128,64,32}B lays it out
template <int RB>
__device__ __forceinline__ int swz(int row) {
  return RB == 128 ? (row & 7) : RB == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1);
}
template <int V>
struct Cfg2<EPI_RESID, V> {
  static constexpr int BN = 256, groups = 4, stages = 4;
  static constexpr int out_bytes = kResidBufs * sub_bytes(EPI_RESID);
  static constexpr int aux_bytes = 4 * 128 * 4;          // per warp group partial sums of squares
};
template <int V>
struct Cfg2<EPI_RESID_DEEP, V> {
  static constexpr int BN = 256, groups = 4, stages = 5;
  static constexpr int out_bytes = kResidBufs * sub_bytes(EPI_RESID_DEEP);
  static constexpr int aux_bytes = 4 * 128 * 4;
};

template <int EPI, int V>
struct Lay {
  using C = Cfg2<EPI, V>;
  static constexpr int BN = C::BN;
  static constexpr int half_n = BN / 2;
  static constexpr int epi_warps = 4 * C::groups;
  static constexpr int threads = (epi_warps + 3) * 32;
  static constexpr int a_bytes = BM * BK * 2;
  static constexpr int b_bytes = half_n * BK * 2;
  static constexpr int stage_bytes = a_bytes + b_bytes;
  static constexpr int ring_bytes = C::stages * stage_bytes;
  static constexpr int out_off = ring_bytes;
  static constexpr int aux_off = out_off + C::out_bytes;
  static constexpr int bar_off = aux_off + C::aux_bytes;
  static constexpr int smem_bytes = bar_off + 512;
  static_assert(stage_bytes % 1024 == 0 && a_bytes % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB shared-memory limit");
};

__device__ __forceinline__ float silu(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

template <int EPI, int V>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Lay<EPI, V>::threads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
             const __grid_constant__ CUtensorMap tm2, const GemmArgs p) {
  using L = Lay<EPI, V>;
  using C = Cfg2<EPI, V>;
  constexpr int BN = L::BN;
  constexpr int E = L::epi_warps;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // no static shared memory in this kernel: the window starts aligned

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* empty_bar = full_bar + C::stages;
  uint64_t* acc_full = empty_bar + C::stages;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* store_ready = acc_empty + 2;  // [4] (one per residual buffer; [0] otherwise)
  uint64_t* out_free = store_ready + 4;   // [2]
  uint64_t* aux_full = out_free + 2;      // [4] residual piece / rotary tables landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + 4);
  uint8_t* out_base = smem + L::out_off;
  uint8_t* aux_base = smem + L::aux_off;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_n = (p.N + BN - 1) / BN;   // only EPI_LOGPROB may have a ragged last column block (masked)
  const int num_pair_m = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_tiles = num_pair_m * num_n;
  const int num_k_blocks = (p.K + BK - 1) / BK;
  // Clusters interleave over the tiles, n fastest: the 74 co-resident pairs then work on ~5 row blocks x all column
  // blocks at a time, so each A box is wanted by up to num_n pairs and each W box by ~5 within a short window and L2
  // serves them from one read (the L2 slices cap at ~6.3 KB/clk chip-wide; 74 pairs x 62 KB per k-block un-merged is
  // 730 cycles against 480-512 of MMA).  QKV used to walk a contiguous range per pair (one row block at a time, a
  // left-over of the per-row cos/sin loads): every pair then streamed its own A rows and the ring ran dry.
  const int tile_begin = cluster_id;
  const int tile_end = num_tiles;
  const int tile_step = num_clusters;
  // row block of a tile: ascending, or (GEMM_REVERSE_M) descending — consecutive launches of the tower alternate, so
  // every launch starts on the rows its producer wrote last (the likeliest to be in L2 still) and ends on the rows its
  // consumer will start with
  const int m_flip = (p.flags & GEMM_REVERSE_M) != 0 ? num_pair_m - 1 : -1;
  auto tile_row = [&](int tile) { return m_flip >= 0 ? m_flip - tile / num_n : tile / num_n; };

  constexpr int kProducerWarp = E, kMmaWarp = E + 1, kStoreWarp = E + 2;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tm0);
    if constexpr (is_resid<EPI>) tma_prefetch_desc(&tm1);
    for (int s = 0; s < C::stages; ++s) {
      mbar_init(&full_bar[s], 1);    // leader's: its own expect_tx arrive; both CTAs' bytes complete on it
      mbar_init(&empty_bar[s], 1);   // multicast tcgen05.commit arrives once in each CTA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * E);  // leader's: every epilogue warp of both CTAs
      mbar_init(&out_free[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&store_ready[s], is_resid<EPI> ? E / kResidBufs : E);
      mbar_init(&aux_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation of both CTAs are visible to the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------ TMA producer (whole warp, one elected lane issues)
    uint32_t it = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      const int m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = (tile % num_n) * BN + static_cast<int>(rank) * L::half_n;
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const uint32_t s = it % C::stages;
        const uint32_t ph = (it / C::stages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* sa = smem + s * L::stage_bytes;
        uint8_t* sb = sa + L::a_bytes;
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);   // the leader's barrier tracks the pair
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full_bar[s], 2 * L::stage_bytes);
          tma_load_2d_2cta(sa, &tmA, bar, kb * BK, m0);
          tma_load_2d_2cta(sb, &tmB, bar, kb * BK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer (leader CTA; whole warp, one elected lane)
    if (leader) {
      const uint32_t idesc = umma_idesc_fmt(umma_idesc_bf16(2 * BM, BN), (p.flags & GEMM_A_F16) != 0, (p.flags & GEMM_A_F16) != 0);
      uint32_t it = 0;
      uint32_t local = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
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
          const uint32_t a_addr = smem_u32(smem + s * L::stage_bytes);
          const uint32_t b_addr = a_addr + L::a_bytes;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t da = umma_desc_sw128(a_addr + k * UK * 2);
              const uint64_t db = umma_desc_sw128(b_addr + k * UK * 2);
              umma_bf16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2cta(&empty_bar[s], 0x3);   // frees this stage in both CTAs
            if (kb == num_k_blocks - 1) umma_commit_2cta(&acc_full[as], 0x3);   // accumulators of both CTAs complete
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kStoreWarp) {
    // ------------------------------------------------------------ store warp: everything the epilogue moves by TMA.
    // The whole warp runs the loops (warp-uniform operands stay in uniform registers); one fixed elected lane issues
    // the bulk copies and owns their bulk groups.
    const bool issuer = elect_one();
    uint32_t local = 0;
    if constexpr (is_resid<EPI>) {
      constexpr int kSubSteps = sub_steps(EPI), kSubCols = sub_cols(EPI), kSubX = sub_x_bytes(EPI), kSubBytes = sub_bytes(EPI);
      // Sub-step st = steps * (tile ordinal) + k covers tile columns [cols * k, +cols) and uses buffer k & 3.  Per buffer:
      //   load residual piece -> epilogue (in place) -> store x, xb -> stores have read the buffer -> load its next piece.
      constexpr int NB = kResidBufs;
      float* red = reinterpret_cast<float*>(aux_base);
      const int n_local = tile_begin < tile_end ? (tile_end - tile_begin + tile_step - 1) / tile_step : 0;
      const int n_sub = kSubSteps * n_local;
      auto coords = [&](int st, int& m0, int& n0, int& n_blk) {
        const int tile = tile_begin + (st / kSubSteps) * tile_step;
        m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
        n_blk = tile % num_n;
        n0 = n_blk * BN + (st % kSubSteps) * kSubCols;
      };
      auto load_x = [&](int st) {   // fp32 residual piece of sub-step st -> its buffer
        int m0, n0, n_blk;
        coords(st, m0, n0, n_blk);
        if (issuer) {
          mbar_expect_tx(&aux_full[st % NB], kSubX);
          tma_load_2d(out_base + (st % NB) * kSubBytes, &tm0, &aux_full[st % NB], n0, m0);
        }
      };
      for (int st = 0; st < NB && st < n_sub; ++st) load_x(st);
      for (int st = 0; st < n_sub; ++st) {
        int m0, n0, n_blk;
        coords(st, m0, n0, n_blk);
        const int b = st % NB;
        mbar_wait(&store_ready[b], (st / NB) & 1u);
        if (issuer) {
          const uint8_t* xb = out_base + b * kSubBytes;
          tma_store_2d(&tm0, xb, n0, m0);
          tma_store_2d(&tm1, xb + kSubX, n0, m0);
          tma_store_commit();
        }
        if (st % kSubSteps == kSubSteps - 1) {
          // sum of squares of the new residual rows over this tile's 256 columns: one value per (row, column block),
          // part-major [N/256][M] so producers and consumers both touch it coalesced; fixed order -> deterministic
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = lane + 32 * k;
            const float sum = (red[row] + red[128 + row]) + (red[256 + row] + red[384 + row]);
            if (m0 + row < p.M) p.ssq[static_cast<int64_t>(n_blk) * p.M + m0 + row] = sum;
          }
        }
        if (st >= 1) {
          // refill one sub-step behind: the previous sub-step's stores have had a whole epilogue step to drain
          if (issuer) tma_store_wait_read<1>();
          if (st - 1 + NB < n_sub) load_x(st - 1 + NB);
        }
        __syncwarp();
      }
    } else if constexpr (EPI == EPI_QKV) {
      // the per-position rotary table (cos | sin of p * f_k, 160 bytes per position) goes to shared memory once
      if (issuer) {
        const uint32_t bytes = static_cast<uint32_t>(p.n_pos) * 160u;
        mbar_expect_tx(&aux_full[0], bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :
                     : "r"(smem_u32(aux_base)), "l"(reinterpret_cast<uint64_t>(p.pos_tab)), "r"(bytes),
                       "r"(smem_u32(&aux_full[0]))
                     : "memory");
      }
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
        const int m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n0 = (tile % num_n) * BN;
        mbar_wait(&store_ready[0], local & 1u);
        if (issuer) {
#pragma unroll
          for (int b = 0; b < 6; ++b) tma_store_2d(&tm0, out_base + b * kBox, n0 + b * 40, m0);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&out_free[0]);
        }
        __syncwarp();
      }
    } else if constexpr (EPI == EPI_LOGPROB) {
      // combine the 4 column groups' (max, sum-exp) per row in a fixed order and write the block's partial,
      // part-major; out_free[parity] tells the epilogue warps the slot may be rewritten
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
        const int m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n_blk = tile % num_n;
        mbar_wait(&store_ready[0], local & 1u);
        const float2* red = reinterpret_cast<const float2*>(aux_base) + (local & 1u) * 4 * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int row = lane + 32 * k;
          float mx = red[row].x;
#pragma unroll
          for (int g = 1; g < 4; ++g) mx = fmaxf(mx, red[g * 128 + row].x);
          float sum = 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float2 v = red[g * 128 + row];
            sum += v.y * (v.x == -INFINITY ? 0.f : __expf(v.x - mx));
          }
          if (m0 + row < p.M) {
            p.part_max[static_cast<int64_t>(n_blk) * p.M + m0 + row] = mx;
            p.part_sum[static_cast<int64_t>(n_blk) * p.M + m0 + row] = sum;
          }
        }
        __syncwarp();
        if (issuer) mbar_arrive(&out_free[local & 1u]);
      }
    } else if constexpr (EPI == EPI_MROPE) {
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
        const int m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n0 = (tile % num_n) * BN;
        mbar_wait(&store_ready[0], local & 1u);
        if (issuer) {
#pragma unroll
          for (int b = 0; b < 4; ++b) tma_store_2d(&tm0, out_base + b * 16384, n0 + b * 64, m0);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(&out_free[0]);
        }
        __syncwarp();
      }
    } else {  // EPI_SWIGLU
      for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
        const int m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
        const int n_blk = tile % num_n;
        mbar_wait(&store_ready[0], local & 1u);
        if (issuer) {
          tma_store_2d(&tm0, out_base, n_blk * (BN / 2), m0);
          tma_store_2d(&tm0, out_base + 16384, n_blk * (BN / 2) + 64, m0);
          tma_store_commit();
          tma_store_wait_read<0>();              // the store has read the buffer
          mbar_arrive(&out_free[0]);
        }
        __syncwarp();
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before exit
  } else {
    // ------------------------------------------------------------ epilogue warps: thread = accumulator row
    const int quarter = warp & 3;   // TMEM lane quarter this warp may read
    const int cg = warp >> 2;       // column group
    const int row = quarter * 32 + lane;   // row within the CTA's 128
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), 0);
    uint32_t local = 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step, ++local) {
      const int n_blk = tile % num_n;
      const int m0 = tile_row(tile) * (2 * BM) + static_cast<int>(rank) * BM;
      const int n0 = n_blk * BN;
      const uint32_t as = local & 1u;
      const uint32_t use = local >> 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kAccStride;
      const int r = m0 + row;
      const bool valid = r < p.M;

      if constexpr (is_resid<EPI>) {
        constexpr int kSubSteps = sub_steps(EPI), kSubCols = sub_cols(EPI), kSubX = sub_x_bytes(EPI), kSubBytes = sub_bytes(EPI);
        constexpr int XR = kSubCols * 4, BR = kSubCols * 2;   // bytes per fp32 / 16-bit row of a sub-step box
        // (one warp per TMEM lane quarter) owns buffer g and the sub-steps k = g, g + 4, ... of every tile; thread = row
        float* red = reinterpret_cast<float*>(aux_base);
        const int g = cg;
        uint8_t* buf = out_base + g * kSubBytes;
        uint8_t* xrow = buf + row * XR;                // 16-byte chunks, chunk index XOR swz<row bytes>(row)
        uint8_t* xbrow = buf + kSubX + row * BR;
        const int sx = swz<XR>(row), sb = swz<BR>(row);
        constexpr int kPer = kSubSteps / kResidBufs;     // sub-steps of a tile per warp group
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int col = (g + kResidBufs * i) * kSubCols;
          mbar_wait(&aux_full[g], (kPer * local + i) & 1u);   // residual piece landed (the buffer's previous stores have drained)
          if (i == 0) {
            mbar_wait_relaxed(&acc_full[as], use & 1u);
            tc_fence_after();
          }
          uint32_t v[kSubCols];
          if constexpr (kSubCols == 32) tmem_ld32(taddr + col, v); else tmem_ld16(taddr + col, v);
          tmem_ld_wait();
          if (i == kPer - 1) {   // all of this warp's TMEM reads of the tile are complete
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) mbar_arrive(&acc_empty[as]); else mbar_arrive_cluster(acc_empty_leader + as * 8);
            }
          }
          uint32_t pk[kSubCols / 2];
          const float ss_before = ss;
#pragma unroll
          for (int j = 0; j < kSubCols / 4; ++j) {
            float4* xp = reinterpret_cast<float4*>(xrow + ((j ^ sx) << 4));
            const float4 xo = *xp;
            const float4 bi = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col + 4 * j));
            const float4 w = __ldg(reinterpret_cast<const float4*>(p.norm_w + n0 + col + 4 * j));
            float4 f;
            f.x = fmaf(__uint_as_float(v[4 * j]), p.in_scale, bi.x) + xo.x;       // in_scale == 1: the plain sum, exactly
            f.y = fmaf(__uint_as_float(v[4 * j + 1]), p.in_scale, bi.y) + xo.y;
            f.z = fmaf(__uint_as_float(v[4 * j + 2]), p.in_scale, bi.z) + xo.z;
            f.w = fmaf(__uint_as_float(v[4 * j + 3]), p.in_scale, bi.w) + xo.w;
            ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss); ss = fmaf(f.z, f.z, ss); ss = fmaf(f.w, f.w, ss);
            *xp = f;                                                         // new residual, in place
            pk[2 * j] = pack16x2<V>(f.x * w.x, f.y * w.y);                   // next GEMM's operand: x * w_norm
            pk[2 * j + 1] = pack16x2<V>(f.z * w.z, f.w * w.w);
          }
          if constexpr (V == 1) {
            // range guard: |x_c * w_c| <= sqrt(sum x^2) * max|w|, so a sub-step whose sum of squares times max w^2 stays
            // below 65504^2 cannot have saturated; otherwise look for the saturated pattern in the packed halves
            if (p.sat != nullptr && !((ss - ss_before) * p.sat_wmax2 < 4.29e9f)) {
              bool hit = false;
#pragma unroll
              for (int j = 0; j < kSubCols / 2; ++j)
                hit |= (pk[j] & 0x7fffu) >= 0x7bffu || ((pk[j] >> 16) & 0x7fffu) >= 0x7bffu;
              if (hit) atomicAdd(p.sat, 1u);
            }
          }
#pragma unroll
          for (int j = 0; j < kSubCols / 8; ++j)
            *reinterpret_cast<uint4*>(xbrow + ((j ^ sb) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          if (i == kPer - 1) red[g * 128 + row] = ss;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&store_ready[g]);
        }
      } else if constexpr (EPI == EPI_QKV) {
        // one head per column group; rotary pairs (d, d + 40) live in one thread's row
        const int col0 = n0 + cg * 80;
        const bool rot = col0 < 2 * (p.N / 3);   // q and k heads rotate, v heads do not
        float rstd = 0.f;
        if (valid) {
          float s = 0.f;
          for (int q = 0; q < p.ssq_parts; ++q) s += p.ssq_in[static_cast<int64_t>(q) * p.M + r];
          rstd = rsqrtf(s * p.inv_dim + p.eps) * p.in_scale;
        }
        int hw = 0;
        if (valid) hw = __ldg(p.pos_win + r);                       // this row's (h, w) position in its frame
        const uint8_t* tab_h = aux_base + (hw & 0xffff) * 160;      // angles 0..19 follow h, 20..39 follow w
        const uint8_t* tab_w = aux_base + (hw >> 16) * 160;
        if (local == 0) mbar_wait(&aux_full[0], 0);                 // the table has landed
        if (local > 0) mbar_wait(&out_free[0], (local - 1) & 1u);   // the previous tile's stores have drained
        mbar_wait_relaxed(&acc_full[as], use & 1u);
        tc_fence_after();
        uint8_t* out_row = out_base + (cg * 2) * kBox + row * 80;
        float amax = 0.f;   // f16 range guard (V == 1)
#pragma unroll 1
        for (int i0 = 0; i0 < 40; i0 += 8) {
          uint32_t lo[8], hi[8];
          tmem_ld8(taddr + cg * 80 + i0, lo);
          tmem_ld8(taddr + cg * 80 + 40 + i0, hi);
          float cs[8], sn[8];
          if (rot) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int f4 = i0 / 4 + j;              // float4 index within the 40 angles: h part 0..4, w part 5..9
              const uint8_t* t = (f4 < 5 ? tab_h : tab_w) + (f4 % 5) * 16;
              const float4 c4 = *reinterpret_cast<const float4*>(t);
              const float4 s4 = *reinterpret_cast<const float4*>(t + 80);
              cs[4 * j] = c4.x; cs[4 * j + 1] = c4.y; cs[4 * j + 2] = c4.z; cs[4 * j + 3] = c4.w;
              sn[4 * j] = s4.x; sn[4 * j + 1] = s4.y; sn[4 * j + 2] = s4.z; sn[4 * j + 3] = s4.w;
            }
          }
          const float4 ba0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i0));
          const float4 ba1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i0 + 4));
          const float4 bb0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 40 + i0));
          const float4 bb1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 40 + i0 + 4));
          const float ba[8] = {ba0.x, ba0.y, ba0.z, ba0.w, ba1.x, ba1.y, ba1.z, ba1.w};
          const float bb[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
          tmem_ld_wait();
          float a[8], b[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            a[q] = fmaf(__uint_as_float(lo[q]), rstd, ba[q]);
            b[q] = fmaf(__uint_as_float(hi[q]), rstd, bb[q]);
          }
          if (rot) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              // q*cos + rotate_half(q)*sin in fp32 (HF :161-166)
              const float na = a[q] * cs[q] - b[q] * sn[q];
              const float nb = b[q] * cs[q] + a[q] * sn[q];
              a[q] = na;
              b[q] = nb;
            }
          }
          if constexpr (V == 1) {
#pragma unroll
            for (int q = 0; q < 8; ++q) amax = fmaxf(amax, fmaxf(fabsf(a[q]), fabsf(b[q])));
          }
          *reinterpret_cast<uint4*>(out_row + i0 * 2) =
              make_uint4(pack16x2<V>(a[0], a[1]), pack16x2<V>(a[2], a[3]), pack16x2<V>(a[4], a[5]), pack16x2<V>(a[6], a[7]));
          *reinterpret_cast<uint4*>(out_row + kBox + i0 * 2) =
              make_uint4(pack16x2<V>(b[0], b[1]), pack16x2<V>(b[2], b[3]), pack16x2<V>(b[4], b[5]), pack16x2<V>(b[6], b[7]));
        }
        if constexpr (V == 1) {
          if (p.sat != nullptr && valid) sat_guard(amax, p.sat);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&acc_empty[as]); else mbar_arrive_cluster(acc_empty_leader + as * 8);
        }
      } else if constexpr (EPI == EPI_LOGPROB) {
        // group cg owns 64 vocabulary columns of the block; thread = token row
        int64_t tgt = -1;
        if (valid) tgt = p.target[r];
        mbar_wait_relaxed(&acc_full[as], use & 1u);
        tc_fence_after();
        const int c0 = n0 + cg * 64;                  // first vocabulary column of this warp's piece
        float mx = -INFINITY, sum = 0.f;
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
          uint32_t v[32];
          tmem_ld32(taddr + cg * 64 + piece * 32, v);
          tmem_ld_wait();
          const int cbase = c0 + piece * 32;
          float pm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float sv = cbase + i < p.N ? __uint_as_float(v[i]) * p.inv_temp : -INFINITY;   // mask the ragged vocabulary tail
            v[i] = __float_as_uint(sv);
            pm = fmaxf(pm, sv);
          }
          if (tgt >= cbase && tgt < cbase + 32 && tgt < p.N) {   // at most one (block, group, piece) per row holds the target
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) t = (tgt - cbase == i) ? __uint_as_float(v[i]) : t;
            p.tgt_logit[r] = t;
          }
          const float nm = fmaxf(mx, pm);
          float ps = 0.f;
          if (nm != -INFINITY) {
#pragma unroll
            for (int i = 0; i < 32; ++i) ps += __expf(__uint_as_float(v[i]) - nm);
            sum = sum * (mx == -INFINITY ? 0.f : __expf(mx - nm)) + ps;
          }
          mx = nm;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&acc_empty[as]); else mbar_arrive_cluster(acc_empty_leader + as * 8);
        }
        // (max, sum) of this warp's 64 columns -> the item parity's slot; the store warp combines the 4 groups
        if (use > 0) mbar_wait(&out_free[as], (use - 1) & 1u);   // the store warp has consumed this slot's previous item
        reinterpret_cast<float2*>(aux_base)[(as * 4 + cg) * 128 + row] = make_float2(mx, sum);
      } else if constexpr (EPI == EPI_MROPE) {
        // tile = two heads of 128 columns; group cg owns head (cg >> 1) and the 32 rotary pairs (f, f + 64) with
        // f in [32 (cg & 1), +32): both halves of a pair live in this thread's accumulator row
        const int head = cg >> 1, f0 = (cg & 1) * 32;
        const int hc = n0 + head * 128;                 // first output column of the head
        const bool rot = hc < p.rot_cols;               // q and k heads rotate, v heads do not
        if (local > 0) mbar_wait(&out_free[0], (local - 1) & 1u);   // the previous tile's stores have drained
        mbar_wait_relaxed(&acc_full[as], use & 1u);
        tc_fence_after();
        // 128B-swizzled [128 x 64 bf16] boxes: 16-byte chunk index XOR (row & 7), as TMA SWIZZLE_128B expects
        uint8_t* box_lo = out_base + (head * 2) * 16384 + row * 128;       // columns [0, 64) of the head
        uint8_t* box_hi = box_lo + 16384;                                  // columns [64, 128)
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 8) {
          const int f = f0 + i0;
          uint32_t lo[8], hi[8];
          tmem_ld8(taddr + head * 128 + f, lo);
          tmem_ld8(taddr + head * 128 + 64 + f, hi);
          float cs[8], sn[8];
          if (rot && valid) {
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.mrope_cos + static_cast<int64_t>(r) * 64 + f));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.mrope_cos + static_cast<int64_t>(r) * 64 + f + 4));
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.mrope_sin + static_cast<int64_t>(r) * 64 + f));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.mrope_sin + static_cast<int64_t>(r) * 64 + f + 4));
            cs[0] = c0.x; cs[1] = c0.y; cs[2] = c0.z; cs[3] = c0.w; cs[4] = c1.x; cs[5] = c1.y; cs[6] = c1.z; cs[7] = c1.w;
            sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) { cs[q] = 1.f; sn[q] = 0.f; }
          }
          float ba[8], bb[8];
          if (p.bias != nullptr) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.bias + hc + f));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.bias + hc + f + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + hc + 64 + f));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + hc + 64 + f + 4));
            ba[0] = a0.x; ba[1] = a0.y; ba[2] = a0.z; ba[3] = a0.w; ba[4] = a1.x; ba[5] = a1.y; ba[6] = a1.z; ba[7] = a1.w;
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) { ba[q] = 0.f; bb[q] = 0.f; }
          }
          tmem_ld_wait();
          float a[8], b[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float x1 = __uint_as_float(lo[q]) + ba[q], x2 = __uint_as_float(hi[q]) + bb[q];
            a[q] = x1 * cs[q] - x2 * sn[q];      // q * cos + rotate_half(q) * sin, in fp32 (HF :149-167 does it in bf16)
            b[q] = x2 * cs[q] + x1 * sn[q];
          }
          const int chunk = f >> 3;              // 16-byte chunk of the 128-byte box row
          *reinterpret_cast<uint4*>(box_lo + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
          *reinterpret_cast<uint4*>(box_hi + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16x2(b[0], b[1]), pack_bf16x2(b[2], b[3]), pack_bf16x2(b[4], b[5]), pack_bf16x2(b[6], b[7]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&acc_empty[as]); else mbar_arrive_cluster(acc_empty_leader + as * 8);
        }
      } else {  // EPI_SWIGLU
        // tile columns [0,128) are gate rows, [128,256) the matching up rows; group cg owns 32 outputs
        float rstd = 0.f;
        if (valid) {
          float s = 0.f;
          for (int q = 0; q < p.ssq_parts; ++q) s += p.ssq_in[static_cast<int64_t>(q) * p.M + r];
          rstd = rsqrtf(s * p.inv_dim + p.eps) * p.in_scale;
        }
        const float rstd_u = rstd * p.up_scale;   // h = silu(g) * (u * up_scale): stored scaled, the up-bias comes pre-scaled
        if (local > 0) mbar_wait(&out_free[0], (local - 1) & 1u);   // the previous tile's store has drained
        mbar_wait_relaxed(&acc_full[as], use & 1u);
        tc_fence_after();
        // 128B-swizzled [128 x 64 bf16] boxes: 16-byte chunk index XOR (row & 7), as TMA SWIZZLE_128B expects
        uint8_t* orow = out_base + (cg >> 1) * 16384 + row * 128;
        float amax = 0.f;   // f16 range guard (V == 1)
#pragma unroll
        for (int piece = 0; piece < 2; ++piece) {
          const int c = cg * 32 + piece * 16;
          uint32_t g[16], u[16];
          tmem_ld16(taddr + c, g);
          tmem_ld16(taddr + BN / 2 + c, u);
          tmem_ld_wait();
          float h[16];
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + q));
            const float4 bu = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + BN / 2 + c + q));
            h[q] = silu(fmaf(__uint_as_float(g[q]), rstd, bg.x)) * fmaf(__uint_as_float(u[q]), rstd_u, bu.x);
            h[q + 1] = silu(fmaf(__uint_as_float(g[q + 1]), rstd, bg.y)) * fmaf(__uint_as_float(u[q + 1]), rstd_u, bu.y);
            h[q + 2] = silu(fmaf(__uint_as_float(g[q + 2]), rstd, bg.z)) * fmaf(__uint_as_float(u[q + 2]), rstd_u, bu.z);
            h[q + 3] = silu(fmaf(__uint_as_float(g[q + 3]), rstd, bg.w)) * fmaf(__uint_as_float(u[q + 3]), rstd_u, bu.w);
          }
          if constexpr (V == 1) {
#pragma unroll
            for (int q = 0; q < 16; ++q) amax = fmaxf(amax, fabsf(h[q]));
          }
          const int chunk = (cg & 1) * 4 + piece * 2;
          *reinterpret_cast<uint4*>(orow + (((chunk) ^ (row & 7)) << 4)) =
              make_uint4(pack16x2<V>(h[0], h[1]), pack16x2<V>(h[2], h[3]), pack16x2<V>(h[4], h[5]), pack16x2<V>(h[6], h[7]));
          *reinterpret_cast<uint4*>(orow + (((chunk + 1) ^ (row & 7)) << 4)) =
              make_uint4(pack16x2<V>(h[8], h[9]), pack16x2<V>(h[10], h[11]), pack16x2<V>(h[12], h[13]), pack16x2<V>(h[14], h[15]));
        }
        if constexpr (V == 1) {
          if (p.sat != nullptr && valid) sat_guard(amax, p.sat);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&acc_empty[as]); else mbar_arrive_cluster(acc_empty_leader + as * 8);
        }
      }
      if constexpr (!is_resid<EPI>) {
        // this warp's part of the tile is in shared memory: publish it to the async proxy, tell the store warp
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&store_ready[0]);
      }
    }
  }

  // No CTA of the pair may exit (or free TMEM) while the other can still read its shared memory or arrive on its
  // barriers.
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) tmem_dealloc_2cta(tmem_base, kTmemCols);
}

bool encode(CUtensorMap* map, CUtensorMapDataType dt, int elem, const void* base, uint64_t inner, uint64_t rows,
            uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, bool swizzle) {
  return tma_encode_2d(map, static_cast<int>(dt), elem, base, inner, rows, row_stride_bytes, box_inner, box_rows, swizzle);
}

template <int EPI, int V>
cudaError_t launch2(const GemmArgs& a, int num_sms, cudaStream_t stream, const char** err) {
  using L = Lay<EPI, V>;
  CUtensorMap tmA, tmB, tm0, tm1, tm2;
  bool ok = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.A, a.K, a.M, a.lda * 2, BK, BM, true) &&
            encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.W, a.K, a.N, a.ldw * 2, BK, L::half_n, true);
  if (EPI == EPI_QKV) {
    ok = ok && encode(&tm0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.out_bf16, a.N, a.M, a.ld_out * 2, 40, BM, false);
    tm1 = tm0;
    tm2 = tm0;
  } else if (is_resid<EPI>) {
    constexpr int SC = sub_cols(EPI);
    ok = ok && tma_encode_2d_swz(&tm0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, a.x, a.N, a.M, a.ldx * 4, SC, BM, SC * 4) &&
         tma_encode_2d_swz(&tm1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a.out_bf16, a.N, a.M, a.ld_out * 2, SC, BM, SC * 2);
    tm2 = tm1;
  } else if (EPI == EPI_MROPE) {
    ok = ok && encode(&tm0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.out_bf16, a.N, a.M, a.ld_out * 2, 64, BM, true);
    tm1 = tm0;
    tm2 = tm0;
  } else if (EPI == EPI_LOGPROB) {
    tm0 = tmA;
    tm1 = tmA;
    tm2 = tmA;
  } else {
    ok = ok && encode(&tm0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a.out_bf16, a.N / 2, a.M, a.ld_out * 2, 64, BM, true);
    tm1 = tm0;
    tm2 = tm0;
  }
  if (!ok) {
    *err = "cuTensorMapEncodeTiled failed (alignment or driver entry point)";
    return cudaErrorInvalidValue;
  }
  static int max_clusters = 0;  // per template instantiation: co-resident CTA pairs on this device
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<EPI, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms & ~1);
    cfg.blockDim = dim3(L::threads);
    cfg.dynamicSmemBytes = L::smem_bytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<EPI, V>, &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) {
      *err = "gemm2: no CTA pair fits on this device";
      return cudaErrorInvalidValue;
    }
    max_clusters = n < num_sms / 2 ? n : num_sms / 2;
    // experiments only (profiles/README.md): cap the number of co-resident CTA pairs, e.g. to see whether a launch is
    // bound by the SMs it occupies or by what all of them share (L2)
    const char* cap = getenv(is_resid<EPI> ? "VR2_GEMM2_MAX_CLUSTERS_RESID" : "VR2_GEMM2_MAX_CLUSTERS");
    if (cap != nullptr) {
      const int c = atoi(cap);
      if (c > 0 && c < max_clusters) max_clusters = c;
    }
  }
  const int pair_tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * ((a.N + L::BN - 1) / L::BN);
  // as many pairs as keep the number of waves: 320 residual tiles on 74 pairs are 5 waves with 24 pairs idle in the last
  // one; 64 pairs do the same 5 tiles each with less contention for what the launch is bound by (L2): measured -2 % on
  // proj, -1 % on down, and a tensor-bound launch (gate|up: 1728 tiles, 24 waves on 72 or 74 pairs) loses nothing
  int clusters = max_clusters;
  if (pair_tiles < clusters) clusters = pair_tiles;
  const int waves = (pair_tiles + clusters - 1) / clusters;
  clusters = (pair_tiles + waves - 1) / waves;
  gemm2_kernel<EPI, V><<<2 * clusters, L::threads, L::smem_bytes, stream>>>(tmA, tmB, tm0, tm1, tm2, a);
  return cudaGetLastError();
}

}  // namespace

constexpr int kResidDeepK = 2048;

bool gemm2_supports(const GemmArgs& a) {
  if (a.A_lo != nullptr || a.row_map != nullptr) return false;
  switch (a.epilogue) {
    case EPI_QKV: return a.N % 240 == 0 && (a.N / 3) % 80 == 0 && a.pos_win != nullptr && a.pos_tab != nullptr && a.n_pos > 0 && a.n_pos <= kMaxPos;
    case EPI_RESID: return (a.flags & GEMM_ADD_RESID) != 0 && a.N % 256 == 0 && a.bias != nullptr;
    case EPI_SWIGLU: return a.N % 256 == 0;
    case EPI_LOGPROB: return a.target != nullptr && a.part_max != nullptr && a.part_sum != nullptr && a.tgt_logit != nullptr;
    case EPI_MROPE: return a.N % 256 == 0 && a.rot_cols % 128 == 0 && a.mrope_cos != nullptr && a.mrope_sin != nullptr && a.out_bf16 != nullptr;
    default: return false;
  }
}

int gemm2_resid_parts(int n) { return n / 256; }

cudaError_t gemm2_launch(const GemmArgs& a_in, int num_sms, cudaStream_t stream, const char** err) {
  GemmArgs a = a_in;
  if (a.in_scale == 0.f) a.in_scale = 1.f;
  if (a.up_scale == 0.f) a.up_scale = 1.f;
  const bool f16 = (a.flags & GEMM_OUT_F16) != 0;
  switch (a.epilogue) {
    case EPI_QKV: return f16 ? launch2<EPI_QKV, 1>(a, num_sms, stream, err) : launch2<EPI_QKV, 0>(a, num_sms, stream, err);
    case EPI_RESID:
      // long K: the deep-ring shape (16-column sub-steps, 5 stages); short K: wide sub-steps, 4 stages
      if (a.K >= kResidDeepK && !getenv("VR2_RESID_WIDE"))
        return f16 ? launch2<EPI_RESID_DEEP, 1>(a, num_sms, stream, err) : launch2<EPI_RESID_DEEP, 0>(a, num_sms, stream, err);
      return f16 ? launch2<EPI_RESID, 1>(a, num_sms, stream, err) : launch2<EPI_RESID, 0>(a, num_sms, stream, err);
    case EPI_SWIGLU: return f16 ? launch2<EPI_SWIGLU, 1>(a, num_sms, stream, err) : launch2<EPI_SWIGLU, 0>(a, num_sms, stream, err);
    case EPI_LOGPROB: return launch2<EPI_LOGPROB, 0>(a, num_sms, stream, err);
    case EPI_MROPE: return launch2<EPI_MROPE, 0>(a, num_sms, stream, err);
    default: *err = "gemm2: unsupported epilogue"; return cudaErrorInvalidValue;
  }
}

}  // namespace vr2
