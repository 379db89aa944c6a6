// Internal (C++) interface of the tcgen05 GEMM used by every dense layer of the
// tower.  C[M,N] = A[M,K] * W[N,K]^T with A and W bf16, K-major; fp32 accumulate
// in TMEM; the epilogue variant decides what is written.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace vr2 {

enum GemmEpilogue : int {
  EPI_STORE = 0,   // out = act(acc + bias) -> fp32 or bf16, optional destination-row map
  EPI_RESID = 1,   // x = [x +] acc + bias (fp32, in place); xb = bf16(x * norm_w); ssq partials
  EPI_QKV = 2,     // v = acc*rstd + bias; 2-D rotary on q,k heads; bf16 [M,3840]
  EPI_SWIGLU = 3,  // h = silu(g*rstd+bg) * (u*rstd+bu), gate/up interleaved per 256-column tile
  EPI_LOGPROB = 4, // LM head without logits: per (row, 256-column block) max / sum-exp of acc * inv_temp, target logit
  EPI_MROPE = 6,   // decoder q|k|v projection with the multimodal rotary embedding fused (CTA-pair kernel): out =
                   // bf16(rot(acc + bias)), heads of 128 columns, pairs (d, d + 64) rotated by per-row cos / sin tables
                   // for columns < rot_cols (the q and k heads); v heads get the bias only
  EPI_STORE_MULTI = 5,  // out = acc + bias as bf16, each (row-mapped) row written to n_multi destination buffers
                        // (local + peer GPUs' copies of the result: the all-gather fused into the producing GEMM)
};

constexpr int kMaxMultiOut = 8;

enum GemmFlags : int {
  GEMM_ADD_RESID = 1,  // EPI_RESID: read-modify-write x
  GEMM_GELU = 2,       // EPI_STORE: exact-erf GELU
  GEMM_OUT_F32 = 4,    // EPI_STORE: fp32 output (else bf16)
  GEMM_A_F16 = 8,      // A and W hold fp16 (the f16 activation mode); kind::f16 rejects mixed fp16 x bf16 operands
  GEMM_OUT_F16 = 16,   // the 16-bit output (out_bf16: EPI_STORE / xb / qkv / h) is written as fp16, saturating
  GEMM_REVERSE_M = 32, // CTA-pair kernel: walk the row blocks from the last one down.  Same tiles, same arithmetic per
                       // element; a launch that starts where its producer stopped finds those rows in L2 still
};

struct GemmArgs {
  // operands (device pointers, bf16, row-major, leading dimension in elements)
  const __nv_bfloat16* A;
  const __nv_bfloat16* A_lo;   // non-null selects the split ("precise") GEMM: C = (A + A_lo) W^T, same layout as A
  int64_t lda;
  const __nv_bfloat16* W;
  int64_t ldw;
  int M, N, K;
  int epilogue;
  int flags;
  // epilogue operands
  const float* bias;       // [N] fp32 or null
  float* x;                // EPI_RESID: residual stream fp32 [rows, ldx]
  int64_t ldx;
  __nv_bfloat16* out_bf16; // bf16 output (EPI_STORE / xb of EPI_RESID / qkv / h)
  __nv_bfloat16* out_lo;   // split mode: rounding residual of out_bf16 (same layout); optional for EPI_STORE
  float* out_f32;          // EPI_STORE fp32 output
  int64_t ld_out;
  const int* row_map;      // destination row of source row r (null = identity)
  const float* norm_w;     // EPI_RESID: weight of the RMSNorm that reads x next, [N]
  float* ssq;              // partial sum of squares: written part-major [parts][M] by EPI_RESID (parts = gemm_resid_parts())
  const float* ssq_in;     // read by EPI_QKV / EPI_SWIGLU, [ssq_parts][M]
  int ssq_parts;
  float inv_dim;           // 1 / hidden
  float eps;
  const float* cos_tab;    // EPI_QKV: [M, 40] fp32
  const float* sin_tab;
  // EPI_LOGPROB (CTA-pair kernel only): columns >= N are masked; partials are part-major [ceil(N/256)][M]
  const int64_t* target;   // [M] column index whose scaled logit is wanted (out of range: none)
  float inv_temp;          // 1 / temperature
  float* part_max;         // [parts][M] max over the block of acc * inv_temp
  float* part_sum;         // [parts][M] sum of exp(acc * inv_temp - max)
  float* tgt_logit;        // [M] acc[target] * inv_temp
  // EPI_QKV, CTA-pair kernel: the same angles indexed by position (plan.cuh pos_win / pos_tab); optional — without
  // them (or with more than 64 positions) the single-CTA kernel runs with the per-row tables
  // f16 activation mode (GEMM_OUT_F16): device counter of conversions that saturated (null: not counted), and the
  // largest norm_w^2 (EPI_RESID of the pair kernel: a row whose sum of squares times this stays below 65504^2 cannot
  // saturate, so the element-wise scan only runs for rows that might)
  unsigned int* sat;
  float sat_wmax2;
  // f16 mode's range shifts (powers of two; 0 is read as 1): the A operand of this GEMM was stored scaled by
  // 1 / in_scale, so the accumulator is multiplied back (EPI_QKV / EPI_SWIGLU: folded into rstd; EPI_RESID: acc *
  // in_scale + bias + x).  up_scale (EPI_SWIGLU): the `up` half is produced times up_scale, i.e. h is stored scaled,
  // with the caller passing an up-bias that is already multiplied by it — no extra instruction per element.
  float in_scale;
  float up_scale;
  // EPI_MROPE: cos / sin of the 64 rotary angles of every row, fp32 [M, 64] each (vr2_mrope_table), and the number of
  // leading output columns that are rotated
  const float* mrope_cos;
  const float* mrope_sin;
  int rot_cols;
  // EPI_STORE_MULTI: destination buffers (row pitch ld_out each), device or peer-mapped device memory
  void* multi_out[kMaxMultiOut];
  int n_multi;
  const int32_t* pos_win;  // [M] hpos | wpos << 16
  const float* pos_tab;    // [n_pos, 40]
  int n_pos;
};

// Enqueues the GEMM on `stream`.  Returns cudaSuccess or the launch error; shape
// violations set *err_msg (static string) and return cudaErrorInvalidValue.
cudaError_t gemm_launch(const GemmArgs& args, int num_sms, cudaStream_t stream, const char** err_msg);

// Number of sum-of-squares partials per row this EPI_RESID launch writes (depends on the kernel that runs it).
int gemm_resid_parts(const GemmArgs& args);

// CTA-pair (cta_group::2) kernel for the QKV / residual / gate|up GEMMs (gemm2.cu); gemm_launch() routes to it
// whenever gemm2_supports().
bool gemm2_supports(const GemmArgs& args);
int gemm2_resid_parts(int n);
cudaError_t gemm2_launch(const GemmArgs& args, int num_sms, cudaStream_t stream, const char** err_msg);

// dtype: a CUtensorMapDataType; swizzle128: SWIZZLE_128B (else none)
bool tma_encode_2d(CUtensorMap* map, int dtype, int elem_bytes, const void* base, uint64_t inner, uint64_t rows,
                   uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, bool swizzle128);

// swizzle_bytes: 0 (none), 32, 64 or 128
bool tma_encode_2d_swz(CUtensorMap* map, int dtype, const void* base, uint64_t inner, uint64_t rows,
                       uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes);

// Resolves cuTensorMapEncodeTiled through the runtime (no link against libcuda).
bool tma_encode_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                        uint32_t box_inner, uint32_t box_rows);

}  // namespace vr2
