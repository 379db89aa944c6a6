// Persistent, warp-specialised tcgen05 GEMM for sm_100a — the single-CTA kernel.
//
//   C[M,N] = A[M,K] * W[N,K]^T      A, W: bf16, K-major      accumulate: fp32 in TMEM
//
// It runs the GEMMs the CTA-pair kernel (gemm2.cu) does not take: the patch embed (rows scatter through the window
// permute), the merger (EPI_STORE: GELU, inverse permute / caller's scatter fused into the store) and every GEMM of
// the split precision (SPLIT).  gemm_launch() routes the bf16-precision QKV / residual / gate|up GEMMs to gemm2.
//
// One CTA per SM loops over 128 x BN output tiles (tile id -> n fastest, so the
// 148 resident CTAs share a band of A rows and all of W through L2).  Roles:
//   TMA producer: 128B-swizzled A/W k-blocks into a 4-stage ring
//   tcgen05.mma issuer: 4 x (128 x BN x 16) per k-block, two accumulator stages in
//                      TMEM (2 x 256 columns)
//   warps 0..4G-1      epilogue: G column groups x 4 lane quarters (a warp may only
//                      read TMEM lanes 32*(warp%4)..+31); tcgen05.ld of the finished
//                      accumulator (thread = row), fused bias / residual / RMS
//                      scale / rotary / SwiGLU / GELU, stores; overlaps the next
//                      tile's MMAs.  G=4 (16 warps) for BN=256, G=3 (one head each)
//                      for the QKV tile of 240 columns
//   warp 4G (one lane)   TMA producer,  warp 4G+1 (one lane)  tcgen05.mma issuer
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
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages

// SPLIT ("precise" mode): the A operand arrives as two bf16 tensors A_hi + A_lo (lo = the rounding
// residual of hi), both multiplied with the same W k-block into one accumulator: 2x the MMA work,
// ~16 mantissa bits on the activations.
template <int BN, bool SPLIT>
struct Cfg {
  static constexpr int groups = BN == 240 ? 3 : 4;        // epilogue column groups
  static constexpr int epi_warps = 4 * groups;
  static constexpr int threads = (epi_warps + 2) * 32;
  static constexpr int group_cols = BN / groups;           // 64 (BN=256), 80 (BN=240: one head) or 40 (BN=160)
  static constexpr int a_half = BM * BK * 2;               // one A tile (hi or lo)
  static constexpr int a_bytes = a_half * (SPLIT ? 2 : 1);
  static constexpr int b_bytes = BN * BK * 2;
  static constexpr int stage_bytes = a_bytes + b_bytes;
  static constexpr int stages = (196 * 1024) / stage_bytes;
  static constexpr int bar_bytes = 256;
  static constexpr int staging_bytes = epi_warps * 2048;   // per-warp epilogue staging
  static constexpr int smem_bytes = stages * stage_bytes + bar_bytes + staging_bytes + 1024;  // +1024: manual alignment
  static_assert(smem_bytes <= 232448, "exceeds the 227 KB shared-memory limit");
  static_assert(stage_bytes % 1024 == 0, "stage must keep 1024-byte alignment for SWIZZLE_128B");
  static_assert(stages >= 3, "pipeline too shallow");
};

constexpr int kStageBytes = 2048;  // per epilogue warp

// EPI_STORE_MULTI (the merger's last GEMM in the sharded multi-GPU mode): the bf16 tile is staged whole in shared
// memory — [128 rows][256 columns] with 528-byte rows (512 + 16: thread-per-row 16-byte stores hit every bank once) —
// and every row leaves as ONE 512-byte bulk copy per destination buffer.  The destinations are the same rows of
// `n_multi` buffers, the local result and the peers' (NVLink-mapped) copies of it: the all-gather of the merged
// embeddings happens tile by tile inside the GEMM that produces them instead of as a collective behind it.
// The ring gives up one stage for the 66 KB tile.
constexpr int kMultiRowBytes = 528;
constexpr int kMultiTileBytes = BM * kMultiRowBytes;

template <int BN, int EPI, bool SPLIT>
__host__ __device__ constexpr int stages_of() {
  // the split operand's stages are 64 KB (A_hi | A_lo | W): two of them beside the tile (the fallback precision)
  return EPI == EPI_STORE_MULTI ? (SPLIT ? 2 : 3) : Cfg<BN, SPLIT>::stages;
}
template <int BN, int EPI, bool SPLIT>
__host__ __device__ constexpr int smem_of() {
  using C = Cfg<BN, SPLIT>;
  return EPI == EPI_STORE_MULTI ? stages_of<BN, EPI, SPLIT>() * C::stage_bytes + C::bar_bytes + kMultiTileBytes + 1024 : C::smem_bytes;
}

// Per-warp staging tile that turns thread-per-row register data into row-contiguous global
// accesses.  Two views of the same 2 KB: fp32 [32 rows][16] (64-byte rows, 4 x 16-byte chunks,
// chunk index XOR-swizzled by (row >> 1) & 3 so both the row-per-thread and the 8-rows-per-
// instruction patterns are bank-conflict free) and bf16 [32 rows][16] (32-byte rows).
struct Stage {
  uint8_t* base;
  __device__ __forceinline__ float4* f4(int row, int chunk) const {
    return reinterpret_cast<float4*>(base + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4));
  }
  __device__ __forceinline__ uint4* h8(int row, int half) const {
    return reinterpret_cast<uint4*>(base + row * 32 + (half << 4));
  }
  // The same [32 x width] fp32 piece, split so the global loads can be issued one chunk ahead of their use:
  // fetch() -> registers (coalesced: 8 rows x 64 B per instruction), put() -> staging.
  static __device__ __forceinline__ void fetch_f32(float4 (&pre)[4], const float* g, int64_t ld, int rows_valid, int lane,
                                                   int width) {
    const int chunk = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = (lane >> 2) + 8 * k;
      pre[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows_valid && chunk * 4 < width) pre[k] = *reinterpret_cast<const float4*>(g + row * ld + chunk * 4);
    }
  }
  __device__ __forceinline__ void put_f32(const float4 (&pre)[4], int lane) const {
    const int chunk = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) *f4((lane >> 2) + 8 * k, chunk) = pre[k];
  }
  __device__ __forceinline__ void store_f32_w(float* g, int64_t ld, int rows_valid, int lane, int width) const {
    const int chunk = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = (lane >> 2) + 8 * k;
      if (row < rows_valid && chunk * 4 < width) *reinterpret_cast<float4*>(g + row * ld + chunk * 4) = *f4(row, chunk);
    }
  }
  __device__ __forceinline__ void store_f32(float* g, int64_t ld, int rows_valid, int lane) const {
    const int chunk = lane & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = (lane >> 2) + 8 * k;
      if (row < rows_valid) *reinterpret_cast<float4*>(g + row * ld + chunk * 4) = *f4(row, chunk);
    }
  }
  __device__ __forceinline__ void store_bf16_w(__nv_bfloat16* g, int64_t ld, int rows_valid, int lane, int width) const {
    const int half = lane & 1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int row = (lane >> 1) + 16 * k;
      if (row < rows_valid && half * 8 < width) *reinterpret_cast<uint4*>(g + row * ld + half * 8) = *h8(row, half);
    }
  }
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// f16: GEMM_OUT_F16 (the f16 activation mode; never together with SPLIT)
// sat: the f16 mode's saturation counter (null otherwise): bumped when a value outside fp16's finite range is converted
template <bool SPLIT>
__device__ __forceinline__ void pack16(const float* f, uint32_t* hi, uint32_t* lo, bool f16, unsigned int* sat) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if constexpr (SPLIT) split_bf16x2(f[2 * j], f[2 * j + 1], hi[j], lo[j]);
    else hi[j] = pack16x2_rt(f[2 * j], f[2 * j + 1], f16);
  }
  if constexpr (!SPLIT) {
    if (sat != nullptr) {
      float amax = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(f[j]));
      sat_guard(amax, sat);
    }
  }
}
// This thread's row, 16 consecutive columns -> the bf16 piece whose (row0, column) origin is g (and its
// rounding residual to g_lo), row-contiguous through the warp's staging tile.
template <bool SPLIT>
__device__ __forceinline__ void emit_piece(const Stage& st, const float* f, __nv_bfloat16* g, __nv_bfloat16* g_lo,
                                           int64_t ld, int rows_valid, int lane, int width, bool f16, unsigned int* sat) {
  uint32_t hi[8], lo[8];
  pack16<SPLIT>(f, hi, lo, f16, sat);
  *st.h8(lane, 0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *st.h8(lane, 1) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  __syncwarp();
  st.store_bf16_w(g, ld, rows_valid, lane, width);
  __syncwarp();
  if constexpr (SPLIT) {
    *st.h8(lane, 0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *st.h8(lane, 1) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    __syncwarp();
    st.store_bf16_w(g_lo, ld, rows_valid, lane, width);
    __syncwarp();
  }
}
// Same piece, thread-per-row (scattered destination rows): o / o_lo point at this row's 16 columns.
template <bool SPLIT>
__device__ __forceinline__ void emit_row(const float* f, __nv_bfloat16* o, __nv_bfloat16* o_lo, bool f16, unsigned int* sat) {
  uint32_t hi[8], lo[8];
  pack16<SPLIT>(f, hi, lo, f16, sat);
  *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(o + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  if constexpr (SPLIT) {
    if (o_lo != nullptr) {
      *reinterpret_cast<uint4*>(o_lo) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(o_lo + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
  }
}

template <int BN, int EPI, bool SPLIT>
__global__ void __launch_bounds__(Cfg<BN, SPLIT>::threads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
            const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  using C = Cfg<BN, SPLIT>;
  constexpr int kStages = stages_of<BN, EPI, SPLIT>();
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment in the shared window.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::stage_bytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* stage_base = smem + kStages * C::stage_bytes + C::bar_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool of16 = (p.flags & GEMM_OUT_F16) != 0;
  unsigned int* const osat = of16 ? p.sat : nullptr;   // range guard of the f16 activation mode
  const int num_n_blocks = (p.N + BN - 1) / BN;
  const int num_m_blocks = (p.M + BM - 1) / BM;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_k_blocks = (p.K + BK - 1) / BK;

  constexpr int kProducerWarp = C::epi_warps;
  constexpr int kMmaWarp = C::epi_warps + 1;
  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmA);
    if constexpr (SPLIT) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], C::epi_warps);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ------------------------------------------------------------ TMA producer (whole warp; one elected lane issues)
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n_blocks) * BM;
      const int n0 = (tile % num_n_blocks) * BN;
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const uint32_t s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* sa = smem + s * C::stage_bytes;
        uint8_t* sb = sa + C::a_bytes;
        if (elect_one()) {
          mbar_expect_tx(&full_bar[s], C::stage_bytes);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
          if constexpr (SPLIT) tma_load_2d(sa + C::a_half, &tmA2, &full_bar[s], kb * BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer (whole warp; one elected lane issues)
    const uint32_t idesc = umma_idesc_fmt(umma_idesc_bf16(BM, BN), (p.flags & GEMM_A_F16) != 0, (p.flags & GEMM_A_F16) != 0);
    uint32_t it = 0;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const uint32_t as = local & 1u;
      const uint32_t use = local >> 1;
      mbar_wait(&acc_empty[as], (use & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kAccStride;
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const uint32_t s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * C::stage_bytes);
        const uint32_t b_addr = a_addr + C::a_bytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * UK * 2);
            const uint64_t db = umma_desc_sw128(b_addr + k * UK * 2);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (SPLIT) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t da = umma_desc_sw128(a_addr + C::a_half + k * UK * 2);
              const uint64_t db = umma_desc_sw128(b_addr + k * UK * 2);
              umma_bf16(tmem_d, da, db, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
          if (kb == num_k_blocks - 1) umma_commit(&acc_full[as]);    // accumulator complete -> epilogue
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 0 .. 4G-1)
    // thread = accumulator row.  A row-per-thread global access touches 32 cache lines per
    // instruction (half-written sectors); every piece therefore goes through this warp's 2 KB
    // staging buffer so that global loads/stores are row-contiguous (ncu r01b: 32 sectors/request,
    // L1 64 %, tensor pipe 24-54 % before; see profiles/).
    const int quarter = warp & 3;          // TMEM lane quarter this warp may read
    const int cg = warp >> 2;              // column group
    constexpr int GC = C::group_cols;
    const Stage st{stage_base + warp * kStageBytes};
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int n_blk = tile % num_n_blocks;
      const int m0 = (tile / num_n_blocks) * BM;
      const int n0 = n_blk * BN;
      const uint32_t as = local & 1u;
      const uint32_t use = local >> 1;
      if constexpr (EPI != EPI_RESID && EPI != EPI_QKV) {   // those two start their global loads first
        mbar_wait(&acc_full[as], use & 1u);
        tc_fence_after();
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * kAccStride;
      const int row0 = m0 + quarter * 32;    // first row of this warp
      const int r = row0 + lane;
      const bool valid = r < p.M;
      const int rows_valid = min(32, max(0, p.M - row0));

      if constexpr (EPI == EPI_STORE_MULTI) {
        // thread = row: its GC columns -> the staged tile; the quarter's four column-group warps meet on a named
        // barrier, then the cg == 0 warp sends each of its 32 rows to every destination with one bulk copy
        static_assert(BN == 256, "the multi-destination store stages a 128 x 256 tile");
        const int64_t dest = valid ? (p.row_map != nullptr ? p.row_map[r] : r) : -1;
        uint8_t* srow = stage_base + (quarter * 32 + lane) * kMultiRowBytes;
        if (cg == 0) tma_store_wait_read<0>();                   // the previous tile's copies have read this row
        named_bar_sync(1 + quarter, 128);                        // ... and the other column groups may overwrite it
#pragma unroll 1
        for (int c = cg * GC; c < (cg + 1) * GC; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr && n0 + c + j < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
            pk[j / 2] = pack_bf16x2(__uint_as_float(v[j]) + b.x, __uint_as_float(v[j + 1]) + b.y);
            pk[j / 2 + 1] = pack_bf16x2(__uint_as_float(v[j + 2]) + b.z, __uint_as_float(v[j + 3]) + b.w);
          }
          *reinterpret_cast<uint4*>(srow + c * 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(srow + c * 2 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        fence_proxy_async();                                     // this thread's staged columns -> visible to the bulk copies
        named_bar_sync(1 + quarter, 128);
        if (cg == 0 && dest >= 0) {
          const int cols = min(BN, p.N - n0);                    // N % 8 == 0: a multiple of 16 bytes
          for (int d = 0; d < p.n_multi; ++d) {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.multi_out[d]) + dest * p.ld_out + n0;
            bulk_store(o, srow, static_cast<uint32_t>(cols) * 2u);
          }
        }
        if (cg == 0) tma_store_commit();
      } else if constexpr (EPI == EPI_STORE) {
        const bool direct = p.row_map != nullptr;   // scattered rows (inverse window permute): row-per-thread stores
        const int64_t dest = valid ? (direct ? p.row_map[r] : r) : 0;
        const bool store_row = valid && dest >= 0;   // a negative destination = "no place for this token" (scatter)
#pragma unroll 1
        for (int c = cg * GC; c < (cg + 1) * GC; c += 16) {
          if (n0 + c >= p.N) break;        // N % 32 == 0 is checked on the host
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
            f[j] = __uint_as_float(v[j]) + b.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b.w;
          }
          if (p.flags & GEMM_GELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
          }
          if (p.flags & GEMM_OUT_F32) {
            if (direct) {
              if (store_row) {
                float* o = p.out_f32 + dest * p.ld_out + n0 + c;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) *st.f4(lane, j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              __syncwarp();
              st.store_f32(p.out_f32 + static_cast<int64_t>(row0) * p.ld_out + n0 + c, p.ld_out, rows_valid, lane);
              __syncwarp();
            }
          } else {
            const int64_t off = n0 + c;
            if (direct) {
              if (store_row) emit_row<SPLIT>(f, p.out_bf16 + dest * p.ld_out + off, p.out_lo ? p.out_lo + dest * p.ld_out + off : nullptr, of16, osat);
            } else if (SPLIT && p.out_lo == nullptr) {
              emit_piece<false>(st, f, p.out_bf16 + static_cast<int64_t>(row0) * p.ld_out + off, nullptr, p.ld_out, rows_valid, lane, 16, of16, osat);
            } else {
              emit_piece<SPLIT>(st, f, p.out_bf16 + static_cast<int64_t>(row0) * p.ld_out + off,
                                p.out_lo + static_cast<int64_t>(row0) * p.ld_out + off, p.ld_out, rows_valid, lane, 16, of16, osat);
            }
          }
        }
      } else if constexpr (EPI == EPI_RESID) {
        // (the accumulator wait for this variant happens inside, after the first residual piece is in flight)
        const bool direct = p.row_map != nullptr;   // patch embed: rows scatter through the window permute
        const bool resid = (p.flags & GEMM_ADD_RESID) != 0;
        const int64_t dest = valid ? (direct ? p.row_map[r] : r) : 0;
        const int c_begin = cg * GC, c_end = (cg + 1) * GC;
        float* xpiece = p.x + static_cast<int64_t>(row0) * p.ldx + n0;   // this warp's rows, tile column 0
        // The residual read is the long pole of this epilogue (x streams from HBM): the first 16-column piece is
        // requested before the accumulator is even complete and every later piece one iteration ahead of its use
        // (ncu r01c: 65 % of the stall samples sat on the LDG -> STS dependency when it was loaded in place).
        float4 pre[4];
        if (!direct && resid) Stage::fetch_f32(pre, xpiece + c_begin, p.ldx, rows_valid, lane, min(16, c_end - c_begin));
        mbar_wait(&acc_full[as], use & 1u);
        tc_fence_after();
        float ss = 0.f;
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += 16) {
          const int wdt = min(16, c_end - c);     // 16, or 8 for the tail of a 40-column group
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          float* xg = xpiece + c;                  // piece origin (identity rows)
          if (!direct && resid) {
            st.put_f32(pre, lane);
            if (c + 16 < c_end) Stage::fetch_f32(pre, xg + 16, p.ldx, rows_valid, lane, min(16, c_end - c - 16));
          }
          tmem_ld_wait();
          __syncwarp();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 xo = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < wdt) {
              a4 = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              if (p.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + j));
              if (resid) {
                if (direct) { if (valid) xo = *reinterpret_cast<const float4*>(p.x + dest * p.ldx + n0 + c + j); }
                else xo = *st.f4(lane, j >> 2);
              }
            }
            f[j] = fmaf(a4.x, p.in_scale, b.x) + xo.x;       // in_scale == 1: the plain sum, exactly
            f[j + 1] = fmaf(a4.y, p.in_scale, b.y) + xo.y;
            f[j + 2] = fmaf(a4.z, p.in_scale, b.z) + xo.z;
            f[j + 3] = fmaf(a4.w, p.in_scale, b.w) + xo.w;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) ss = fmaf(f[j], f[j], ss);
          float xw[16];   // x * w of the RMSNorm that reads x next: the next GEMM's A operand before rounding
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < wdt) w = __ldg(reinterpret_cast<const float4*>(p.norm_w + n0 + c + j));
            xw[j] = f[j] * w.x;
            xw[j + 1] = f[j + 1] * w.y;
            xw[j + 2] = f[j + 2] * w.z;
            xw[j + 3] = f[j + 3] * w.w;
          }
          if (direct) {
            if (valid) {
              float* xr = p.x + dest * p.ldx + n0 + c;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (j < wdt) *reinterpret_cast<float4*>(xr + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              if (wdt == 16) {
                emit_row<SPLIT>(xw, p.out_bf16 + dest * p.ld_out + n0 + c, p.out_lo ? p.out_lo + dest * p.ld_out + n0 + c : nullptr, of16, osat);
              } else {
                uint32_t hi[8], lo[8];
                pack16<SPLIT>(xw, hi, lo, of16, osat);
                *reinterpret_cast<uint4*>(p.out_bf16 + dest * p.ld_out + n0 + c) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                if constexpr (SPLIT) {
                  if (p.out_lo != nullptr)
                    *reinterpret_cast<uint4*>(p.out_lo + dest * p.ld_out + n0 + c) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
              }
            }
          } else {
            __syncwarp();   // every lane has read its residual row before the tile is overwritten
#pragma unroll
            for (int j = 0; j < 4; ++j) *st.f4(lane, j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            __syncwarp();
            st.store_f32_w(xg, p.ldx, rows_valid, lane, wdt);
            __syncwarp();
            const int64_t off = static_cast<int64_t>(row0) * p.ld_out + n0 + c;
            emit_piece<SPLIT>(st, xw, p.out_bf16 + off, p.out_lo + off, p.ld_out, rows_valid, lane, wdt, of16, osat);
          }
        }
        // one partial per (row, column group), part-major [N/40][M]: deterministic, no atomics, coalesced both ways
        if (valid) p.ssq[static_cast<int64_t>(n_blk * C::groups + cg) * p.M + dest] = ss;
      } else if constexpr (EPI == EPI_QKV) {
        // Tile = 3 heads of 80, one head per column group; rotary pairs (d, d+40) live in one thread's row.
        // A head is processed in 5 pieces of 8 pairs.  Everything that does not depend on the accumulator (the
        // row's sum of squares, the first cos / sin piece) is requested before waiting for it, later cos / sin
        // pieces one iteration ahead.  One staging pass carries cos (chunks 0-1) and sin (chunks 2-3) of a
        // piece; one more carries the rotated first halves (columns i0..) and second halves (40 + i0..).
        const int col0 = n0 + cg * 80;
        const bool rot = col0 < 2 * (p.N / 3);  // q and k heads rotate, v heads do not
        const int chunk = lane & 3;
        const float* tab = (chunk < 2 ? p.cos_tab : p.sin_tab) + static_cast<int64_t>(row0) * 40 + (chunk & 1) * 4;
        auto fetch_cs = [&](float4 (&pre)[4], int i0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int row = (lane >> 2) + 8 * k;
            pre[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows_valid) pre[k] = __ldg(reinterpret_cast<const float4*>(tab + row * 40 + i0));
          }
        };
        float4 pre[4];
        if (rot) fetch_cs(pre, 0);
        float rstd = 0.f;
        if (valid) {
          float s = 0.f;
          for (int q = 0; q < p.ssq_parts; ++q) s += p.ssq_in[static_cast<int64_t>(q) * p.M + r];
          rstd = rsqrtf(s * p.inv_dim + p.eps) * p.in_scale;
        }
        mbar_wait(&acc_full[as], use & 1u);
        tc_fence_after();
        const int64_t og = static_cast<int64_t>(row0) * p.ld_out + col0;   // head origin, this warp's rows
#pragma unroll 1
        for (int i0 = 0; i0 < 40; i0 += 8) {
          uint32_t lo[8], hi[8];
          tmem_ld8(taddr + cg * 80 + i0, lo);
          tmem_ld8(taddr + cg * 80 + 40 + i0, hi);
          float4 c0, c1, s0, s1;
          if (rot) {
            st.put_f32(pre, lane);
            __syncwarp();
            c0 = *st.f4(lane, 0); c1 = *st.f4(lane, 1); s0 = *st.f4(lane, 2); s1 = *st.f4(lane, 3);
            __syncwarp();
            if (i0 + 8 < 40) fetch_cs(pre, i0 + 8);   // flies while this piece is rotated, packed and stored
          }
          const float4 ba0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i0));
          const float4 ba1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i0 + 4));
          const float4 bb0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 40 + i0));
          const float4 bb1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 40 + i0 + 4));
          tmem_ld_wait();
          float a[8], b[8];
          a[0] = fmaf(__uint_as_float(lo[0]), rstd, ba0.x); a[1] = fmaf(__uint_as_float(lo[1]), rstd, ba0.y);
          a[2] = fmaf(__uint_as_float(lo[2]), rstd, ba0.z); a[3] = fmaf(__uint_as_float(lo[3]), rstd, ba0.w);
          a[4] = fmaf(__uint_as_float(lo[4]), rstd, ba1.x); a[5] = fmaf(__uint_as_float(lo[5]), rstd, ba1.y);
          a[6] = fmaf(__uint_as_float(lo[6]), rstd, ba1.z); a[7] = fmaf(__uint_as_float(lo[7]), rstd, ba1.w);
          b[0] = fmaf(__uint_as_float(hi[0]), rstd, bb0.x); b[1] = fmaf(__uint_as_float(hi[1]), rstd, bb0.y);
          b[2] = fmaf(__uint_as_float(hi[2]), rstd, bb0.z); b[3] = fmaf(__uint_as_float(hi[3]), rstd, bb0.w);
          b[4] = fmaf(__uint_as_float(hi[4]), rstd, bb1.x); b[5] = fmaf(__uint_as_float(hi[5]), rstd, bb1.y);
          b[6] = fmaf(__uint_as_float(hi[6]), rstd, bb1.z); b[7] = fmaf(__uint_as_float(hi[7]), rstd, bb1.w);
          if (rot) {
            const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
            const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              // q*cos + rotate_half(q)*sin in fp32 (HF :161-166)
              const float na = a[q] * cs[q] - b[q] * sn[q];
              const float nb = b[q] * cs[q] + a[q] * sn[q];
              a[q] = na;
              b[q] = nb;
            }
          }
          // staging bf16 [32][16]: first halves in columns 0-7 (-> head columns i0..), second halves in 8-15 (-> 40 + i0..)
          uint32_t ah[4], al[4], bh[4], bl[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if constexpr (SPLIT) {
              split_bf16x2(a[2 * q], a[2 * q + 1], ah[q], al[q]);
              split_bf16x2(b[2 * q], b[2 * q + 1], bh[q], bl[q]);
            } else {
              ah[q] = pack16x2_rt(a[2 * q], a[2 * q + 1], of16);
              bh[q] = pack16x2_rt(b[2 * q], b[2 * q + 1], of16);
            }
          }
          if (!SPLIT && osat != nullptr) {
            float amax = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) amax = fmaxf(amax, fmaxf(fabsf(a[q]), fabsf(b[q])));
            sat_guard(amax, osat);
          }
          const int half = lane & 1;
          const int64_t piece = og + (half ? 40 + i0 : i0);
          *st.h8(lane, 0) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
          *st.h8(lane, 1) = make_uint4(bh[0], bh[1], bh[2], bh[3]);
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int row = (lane >> 1) + 16 * k;
            if (row < rows_valid) *reinterpret_cast<uint4*>(p.out_bf16 + piece + row * p.ld_out) = *st.h8(row, half);
          }
          __syncwarp();
          if constexpr (SPLIT) {
            *st.h8(lane, 0) = make_uint4(al[0], al[1], al[2], al[3]);
            *st.h8(lane, 1) = make_uint4(bl[0], bl[1], bl[2], bl[3]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int row = (lane >> 1) + 16 * k;
              if (row < rows_valid) *reinterpret_cast<uint4*>(p.out_lo + piece + row * p.ld_out) = *st.h8(row, half);
            }
            __syncwarp();
          }
        }
      } else if constexpr (EPI == EPI_SWIGLU) {
        // Tile columns [0,128) are gate rows, [128,256) the matching up rows; group cg owns 32 outputs.
        float rstd = 0.f;
        if (valid) {
          float s = 0.f;
          for (int q = 0; q < p.ssq_parts; ++q) s += p.ssq_in[static_cast<int64_t>(q) * p.M + r];
          rstd = rsqrtf(s * p.inv_dim + p.eps) * p.in_scale;
        }
        constexpr int OC = BN / 2 / C::groups;  // 32 output columns per group
#pragma unroll 1
        for (int c = cg * OC; c < (cg + 1) * OC; c += 16) {
          uint32_t g[16], u[16];
          tmem_ld16(taddr + c, g);
          tmem_ld16(taddr + BN / 2 + c, u);
          tmem_ld_wait();
          float h[16];
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 bg = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c + q));
            const float4 bu = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + BN / 2 + c + q));
            h[q] = silu(fmaf(__uint_as_float(g[q]), rstd, bg.x)) * fmaf(__uint_as_float(u[q]), rstd * p.up_scale, bu.x);
            h[q + 1] = silu(fmaf(__uint_as_float(g[q + 1]), rstd, bg.y)) * fmaf(__uint_as_float(u[q + 1]), rstd * p.up_scale, bu.y);
            h[q + 2] = silu(fmaf(__uint_as_float(g[q + 2]), rstd, bg.z)) * fmaf(__uint_as_float(u[q + 2]), rstd * p.up_scale, bu.z);
            h[q + 3] = silu(fmaf(__uint_as_float(g[q + 3]), rstd, bg.w)) * fmaf(__uint_as_float(u[q + 3]), rstd * p.up_scale, bu.w);
          }
          const int64_t off = static_cast<int64_t>(row0) * p.ld_out + n_blk * (BN / 2) + c;
          emit_piece<SPLIT>(st, h, p.out_bf16 + off, p.out_lo + off, p.ld_out, rows_valid, lane, 16, of16, osat);
        }
      }
      // all of this warp's TMEM reads are complete (tcgen05.wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
    if constexpr (EPI == EPI_STORE_MULTI) {
      // every bulk copy of this thread has been written (not only read from shared memory), and the writes — some of
      // them into peer GPUs' memory — are ordered before whatever signals the peers after this kernel
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __threadfence_system();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
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

template <int BN, int EPI, bool SPLIT>
cudaError_t launch(const GemmArgs& a, int num_sms, cudaStream_t stream, const char** err) {
  using C = Cfg<BN, SPLIT>;
  CUtensorMap tmA, tmA2, tmB;
  if (!tma_encode_2d_bf16(&tmA, a.A, a.K, a.M, a.lda * 2, BK, BM) ||
      !tma_encode_2d_bf16(&tmA2, SPLIT ? a.A_lo : a.A, a.K, a.M, a.lda * 2, BK, BM) ||
      !tma_encode_2d_bf16(&tmB, a.W, a.K, a.N, a.ldw * 2, BK, BN)) {
    *err = "cuTensorMapEncodeTiled failed (alignment or driver entry point)";
    return cudaErrorInvalidValue;
  }
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, EPI, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_of<BN, EPI, SPLIT>());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_kernel<BN, EPI, SPLIT><<<grid, C::threads, smem_of<BN, EPI, SPLIT>(), stream>>>(tmA, tmA2, tmB, a);
  return cudaGetLastError();
}

template <int BN, int EPI>
cudaError_t launch_any(const GemmArgs& a, int num_sms, cudaStream_t stream, const char** err) {
  return a.A_lo != nullptr ? launch<BN, EPI, true>(a, num_sms, stream, err)
                           : launch<BN, EPI, false>(a, num_sms, stream, err);
}

}  // namespace

bool tma_encode_2d_swz(CUtensorMap* map, int dtype, const void* base, uint64_t inner, uint64_t rows,
                       uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes) {
  EncodeFn fn = resolve_encode();
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                     : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, static_cast<CUtensorMapDataType>(dtype), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tma_encode_2d(CUtensorMap* map, int dtype, int elem_bytes, const void* base, uint64_t inner, uint64_t rows,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, bool swizzle128) {
  (void)elem_bytes;
  return tma_encode_2d_swz(map, dtype, base, inner, rows, row_stride_bytes, box_inner, box_rows, swizzle128 ? 128 : 0);
}

bool tma_encode_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_rows) {
  return tma_encode_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, inner, rows, row_stride_bytes, box_inner, box_rows,
                       true);
}

int gemm_resid_parts(const GemmArgs& a) { return gemm2_supports(a) ? gemm2_resid_parts(a.N) : a.N / 40; }

cudaError_t gemm_launch(const GemmArgs& a_in, int num_sms, cudaStream_t stream, const char** err) {
  GemmArgs a = a_in;
  if (a.in_scale == 0.f) a.in_scale = 1.f;
  if (a.up_scale == 0.f) a.up_scale = 1.f;
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
  if (gemm2_supports(a)) {
    switch (a.epilogue) {
      case EPI_QKV:
        if (a.bias == nullptr || a.ssq_in == nullptr || a.out_bf16 == nullptr) {
          *err = "gemm: EPI_QKV needs bias/ssq/out";
          return cudaErrorInvalidValue;
        }
        break;
      case EPI_RESID:
        if (a.x == nullptr || a.out_bf16 == nullptr || a.norm_w == nullptr || a.ssq == nullptr) {
          *err = "gemm: EPI_RESID needs x/xb/norm_w/ssq";
          return cudaErrorInvalidValue;
        }
        break;
      case EPI_LOGPROB:
      case EPI_MROPE:
        break;
      default:
        if (a.bias == nullptr || a.ssq_in == nullptr || a.out_bf16 == nullptr) {
          *err = "gemm: EPI_SWIGLU needs bias/ssq/out";
          return cudaErrorInvalidValue;
        }
    }
    return gemm2_launch(a, num_sms, stream, err);
  }
  if (a.A_lo != nullptr) {
    if (reinterpret_cast<uintptr_t>(a.A_lo) & 15) {
      *err = "gemm: A_lo must be 16-byte aligned for TMA";
      return cudaErrorInvalidValue;
    }
    if (a.epilogue != EPI_STORE && a.epilogue != EPI_STORE_MULTI && a.out_lo == nullptr) {
      *err = "gemm: split (A_hi + A_lo) GEMM needs out_lo for this epilogue";
      return cudaErrorInvalidValue;
    }
  }
  switch (a.epilogue) {
    case EPI_STORE_MULTI:
      if (a.n_multi < 1 || a.n_multi > kMaxMultiOut || (a.flags & (GEMM_OUT_F32 | GEMM_OUT_F16 | GEMM_GELU)) ||
          a.N % 8 != 0 || (a.ld_out * 2) % 16 != 0) {
        *err = "gemm: EPI_STORE_MULTI takes 1..8 bf16 destinations, N % 8 == 0, no GELU / fp32 output";
        return cudaErrorInvalidValue;
      }
      for (int d = 0; d < a.n_multi; ++d)
        if (a.multi_out[d] == nullptr || (reinterpret_cast<uintptr_t>(a.multi_out[d]) & 15)) {
          *err = "gemm: EPI_STORE_MULTI destinations must be non-null and 16-byte aligned";
          return cudaErrorInvalidValue;
        }
      return launch_any<256, EPI_STORE_MULTI>(a, num_sms, stream, err);
    case EPI_STORE:
      if ((a.flags & GEMM_OUT_F32) ? a.out_f32 == nullptr : a.out_bf16 == nullptr) {
        *err = "gemm: missing output";
        return cudaErrorInvalidValue;
      }
      if (a.N % 32 != 0) {
        *err = "gemm: EPI_STORE needs N % 32 == 0";
        return cudaErrorInvalidValue;
      }
      return launch_any<256, EPI_STORE>(a, num_sms, stream, err);
    case EPI_RESID:
      if (a.N % 160 != 0 || a.x == nullptr || a.out_bf16 == nullptr || a.norm_w == nullptr || a.ssq == nullptr) {
        *err = "gemm: EPI_RESID needs N % 160 == 0 and x/xb/norm_w/ssq";
        return cudaErrorInvalidValue;
      }
      // 128 x 160 tiles: N = 1280 gives 8 column blocks, so M = 16384 is 1024 tiles = 6.92 waves of 148 CTAs
      // (128 x 256 tiles: 640 tiles = 4.32 waves, 13.5 % of the last wave idle)
      return launch_any<160, EPI_RESID>(a, num_sms, stream, err);
    case EPI_QKV:
      if (a.N % 240 != 0 || a.bias == nullptr || a.ssq_in == nullptr || a.cos_tab == nullptr ||
          a.sin_tab == nullptr || a.out_bf16 == nullptr) {
        *err = "gemm: EPI_QKV needs N % 240 == 0 and bias/ssq/cos/sin/out";
        return cudaErrorInvalidValue;
      }
      return launch_any<240, EPI_QKV>(a, num_sms, stream, err);
    case EPI_SWIGLU:
      if (a.N % 256 != 0 || a.bias == nullptr || a.ssq_in == nullptr || a.out_bf16 == nullptr) {
        *err = "gemm: EPI_SWIGLU needs N % 256 == 0 and bias/ssq/out";
        return cudaErrorInvalidValue;
      }
      return launch_any<256, EPI_SWIGLU>(a, num_sms, stream, err);
    case EPI_LOGPROB:
      *err = "gemm: EPI_LOGPROB needs target / partial buffers (CTA-pair kernel only)";
      return cudaErrorInvalidValue;
    case EPI_MROPE:
      *err = "gemm: EPI_MROPE needs N % 256 == 0, rot_cols % 128 == 0, cos / sin tables and an output (CTA-pair kernel only)";
      return cudaErrorInvalidValue;
    default:
      *err = "gemm: unknown epilogue";
      return cudaErrorInvalidValue;
  }
}

}  // namespace vr2
