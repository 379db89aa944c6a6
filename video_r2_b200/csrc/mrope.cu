// Decoder-side M-RoPE application: the consumer of get_rope_index's [3, B, S] position ids.
// Hugging Face computes, once per forward (modeling_qwen2_5_vl.py Qwen2_5_VLRotaryEmbedding.forward),
//     freqs[a, b, s, i] = inv_freq[i] * float(pos[a, b, s])            (fp32; a = temporal / height / width)
//     cos = cat(freqs, freqs).cos() * attention_scaling -> x.dtype     (sin likewise)
// and in every decoder layer (apply_multimodal_rotary_pos_emb, mrope_section = [16, 24, 24] for head_dim 128)
//     channel c of cos/sin takes axis  chunk(c) % 3  over the chunks [16, 24, 24, 16, 24, 24]
//     q_embed = q * cos + rotate_half(q) * sin,   rotate_half(x) = cat(-x[D/2:], x[:D/2])          (k likewise)
// in the tensors' dtype (bf16 in Video-R2: each product and the sum are rounded to bf16).  The reference runs this
// as ~12 elementwise launches per layer, or Liger's Triton kernel (train_sft.py:83-84).
//
// One kernel: cos/sin are formed from the position ids on the fly (never materialised), and applied to every head of
// q and k of a token by the CTA that owns the token.  The bf16 path reproduces HF's roundings operation by operation.
// HBM-bound: algorithmic bytes = read + write of q and k.
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace vr2 {

namespace {

struct MropeArgs {
  int B, S, Hq, Hk, D;
  int sec0, sec1;        // channels [0, sec0) follow axis 0, [sec0, sec0 + sec1) axis 1, the rest of D/2 axis 2
  float scaling;
};

__device__ __forceinline__ float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// grid: (S tokens, B); block: D/2 threads x TOK tokens... here one block = one token, threads = heads x D/2 pairs
template <typename T>
__global__ void mrope_kernel(const T* __restrict__ q, const T* __restrict__ k, T* __restrict__ qo, T* __restrict__ ko,
                             const int64_t* __restrict__ pos, const float* __restrict__ inv_freq, const MropeArgs a) {
  const int s = blockIdx.x, b = blockIdx.y;
  const int half = a.D / 2;
  const int i = threadIdx.x % half;                 // frequency index = channel pair (i, i + D/2)
  const int axis = i < a.sec0 ? 0 : (i < a.sec0 + a.sec1 ? 1 : 2);
  const int64_t p = pos[(static_cast<int64_t>(axis) * a.B + b) * a.S + s];
  const float ang = __fmul_rn(inv_freq[i], static_cast<float>(p));   // inv_freq @ position: one fp32 product
  float c = cosf(ang) * a.scaling, sn = sinf(ang) * a.scaling;
  constexpr bool kBf16 = sizeof(T) == 2;
  if (kBf16) { c = bf16r(c); sn = bf16r(sn); }                        // cos.to(x.dtype), sin.to(x.dtype)
  const int heads = a.Hq + a.Hk;
  for (int h = threadIdx.x / half; h < heads; h += blockDim.x / half) {
    const bool isq = h < a.Hq;
    const int hh = isq ? h : h - a.Hq;
    const int64_t off = ((static_cast<int64_t>(b) * (isq ? a.Hq : a.Hk) + hh) * a.S + s) * a.D;   // [B, H, S, D]
    const T* src = (isq ? q : k) + off;
    T* dst = (isq ? qo : ko) + off;
    float x1, x2;
    if constexpr (kBf16) { x1 = __bfloat162float(src[i]); x2 = __bfloat162float(src[i + half]); }
    else { x1 = src[i]; x2 = src[i + half]; }
    // channel i:        x1 * cos + (-x2) * sin        channel i + D/2:  x2 * cos + x1 * sin
    float o1, o2;
    if constexpr (kBf16) {
      o1 = bf16r(bf16r(x1 * c) + bf16r(-x2 * sn));
      o2 = bf16r(bf16r(x2 * c) + bf16r(x1 * sn));
      dst[i] = __float2bfloat16_rn(o1);
      dst[i + half] = __float2bfloat16_rn(o2);
    } else {
      o1 = __fadd_rn(__fmul_rn(x1, c), __fmul_rn(-x2, sn));          // no FMA contraction: torch multiplies, then adds
      o2 = __fadd_rn(__fmul_rn(x2, c), __fmul_rn(x1, sn));
      dst[i] = o1;
      dst[i + half] = o2;
    }
  }
}

// cos / sin of the D/2 rotary angles of every token, fp32 [B*S, D/2]: what Qwen2_5_VLRotaryEmbedding.forward computes
// once per decoder forward (the same for all 28 layers), with the mrope_section axis selection already applied.
__global__ void mrope_table_kernel(const int64_t* __restrict__ pos, const float* __restrict__ inv_freq, int B, int S, int half,
                                   int sec0, int sec1, float scaling, float* __restrict__ cos_out, float* __restrict__ sin_out) {
  const int64_t total = static_cast<int64_t>(B) * S * half;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(e % half);
    const int64_t tok = e / half;                       // b * S + s
    const int axis = i < sec0 ? 0 : (i < sec0 + sec1 ? 1 : 2);
    const int64_t p = pos[static_cast<int64_t>(axis) * B * S + tok];
    const float ang = __fmul_rn(inv_freq[i], static_cast<float>(p));
    cos_out[e] = cosf(ang) * scaling;
    sin_out[e] = sinf(ang) * scaling;
  }
}

}  // namespace
}  // namespace vr2

using namespace vr2;

extern "C" int vr2_mrope_table(const int64_t* position_ids_dev, int32_t B, int32_t S, const float* inv_freq_dev,
                               int32_t half_dim, const int32_t* mrope_section_host, float attention_scaling,
                               float* cos_dev, float* sin_dev, void* stream_v) {
  if (position_ids_dev == nullptr || inv_freq_dev == nullptr || mrope_section_host == nullptr || cos_dev == nullptr ||
      sin_dev == nullptr)
    return fail(VR2_ERR_ARG, "vr2_mrope_table: null pointer");
  if (B <= 0 || S <= 0 || half_dim <= 0) return fail(VR2_ERR_ARG, "vr2_mrope_table: bad shape B=%d S=%d half_dim=%d", B, S, half_dim);
  if (mrope_section_host[0] < 0 || mrope_section_host[1] < 0 || mrope_section_host[2] < 0 ||
      mrope_section_host[0] + mrope_section_host[1] + mrope_section_host[2] != half_dim)
    return fail(VR2_ERR_ARG, "vr2_mrope_table: mrope_section must sum to head_dim / 2 = %d", half_dim);
  const int64_t total = static_cast<int64_t>(B) * S * half_dim;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
  mrope_table_kernel<<<blocks, 256, 0, as_stream(stream_v)>>>(position_ids_dev, inv_freq_dev, B, S, half_dim,
                                                              mrope_section_host[0], mrope_section_host[1], attention_scaling,
                                                              cos_dev, sin_dev);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}

extern "C" int vr2_decoder_qkv_mrope(const void* hidden_dev, const void* wqkv_dev, const float* bias_dev, int64_t M, int32_t N,
                                     int32_t K, int32_t head_dim, int32_t rot_cols, const float* cos_dev, const float* sin_dev,
                                     void* out_dev, void* stream_v) {
  if (hidden_dev == nullptr || wqkv_dev == nullptr || cos_dev == nullptr || sin_dev == nullptr || out_dev == nullptr)
    return fail(VR2_ERR_ARG, "vr2_decoder_qkv_mrope: null pointer");
  if (head_dim != 128) return fail(VR2_ERR_ARG, "vr2_decoder_qkv_mrope: head_dim %d unsupported (the epilogue is built for 128)", head_dim);
  if (M <= 0 || M > 0x7fffffff || N <= 0 || N % 256 != 0 || K <= 0 || K % 8 != 0 || rot_cols < 0 || rot_cols > N || rot_cols % 128 != 0)
    return fail(VR2_ERR_ARG, "vr2_decoder_qkv_mrope: bad shape M=%lld N=%d K=%d rot_cols=%d (N %% 256, K %% 8, rot_cols %% 128)",
                (long long)M, N, K, rot_cols);
  int dev = 0, sms = 148;
  VR2_CUDA(cudaGetDevice(&dev));
  VR2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GemmArgs a{};
  a.A = static_cast<const __nv_bfloat16*>(hidden_dev); a.lda = K;
  a.W = static_cast<const __nv_bfloat16*>(wqkv_dev); a.ldw = K;
  a.M = static_cast<int>(M); a.N = N; a.K = K;
  a.epilogue = EPI_MROPE;
  a.bias = bias_dev;
  a.out_bf16 = static_cast<__nv_bfloat16*>(out_dev); a.ld_out = N;
  a.mrope_cos = cos_dev; a.mrope_sin = sin_dev; a.rot_cols = rot_cols;
  const char* msg = "";
  cudaError_t e = gemm_launch(a, sms, as_stream(stream_v), &msg);
  if (e != cudaSuccess)
    return fail(e == cudaErrorInvalidValue ? VR2_ERR_ARG : VR2_ERR_CUDA, "vr2_decoder_qkv_mrope: %s %s", cudaGetErrorString(e), msg);
  return VR2_OK;
}

extern "C" int vr2_mrope_apply(const void* q_dev, const void* k_dev, void* q_out_dev, void* k_out_dev, int32_t dtype,
                               int32_t B, int32_t S, int32_t Hq, int32_t Hk, int32_t D, const int64_t* position_ids_dev,
                               const float* inv_freq_dev, const int32_t* mrope_section_host, float attention_scaling,
                               void* stream_v) {
  if (q_dev == nullptr || k_dev == nullptr || q_out_dev == nullptr || k_out_dev == nullptr || position_ids_dev == nullptr ||
      inv_freq_dev == nullptr || mrope_section_host == nullptr)
    return fail(VR2_ERR_ARG, "vr2_mrope_apply: null pointer");
  if (dtype != VR2_F32 && dtype != VR2_BF16) return fail(VR2_ERR_ARG, "vr2_mrope_apply: dtype must be fp32 or bf16");
  if (B <= 0 || S <= 0 || Hq <= 0 || Hk <= 0 || D <= 0 || D % 2 != 0 || D / 2 > 512)
    return fail(VR2_ERR_ARG, "vr2_mrope_apply: bad shape B=%d S=%d Hq=%d Hk=%d D=%d", B, S, Hq, Hk, D);
  if (mrope_section_host[0] < 0 || mrope_section_host[1] < 0 || mrope_section_host[2] < 0 ||
      mrope_section_host[0] + mrope_section_host[1] + mrope_section_host[2] != D / 2)
    return fail(VR2_ERR_ARG, "vr2_mrope_apply: mrope_section (%d,%d,%d) must sum to head_dim/2 = %d", mrope_section_host[0],
                mrope_section_host[1], mrope_section_host[2], D / 2);
  if (B > 65535) return fail(VR2_ERR_ARG, "vr2_mrope_apply: batch %d too large", B);
  MropeArgs a{B, S, Hq, Hk, D, mrope_section_host[0], mrope_section_host[1], attention_scaling};
  const int half = D / 2;
  int hpb = 512 / half;                         // heads processed concurrently by a block
  if (hpb < 1) hpb = 1;
  if (hpb > Hq + Hk) hpb = Hq + Hk;
  dim3 grid(S, B), block(half * hpb);
  cudaStream_t s = as_stream(stream_v);
  if (dtype == VR2_BF16)
    mrope_kernel<<<grid, block, 0, s>>>(static_cast<const __nv_bfloat16*>(q_dev), static_cast<const __nv_bfloat16*>(k_dev),
                                        static_cast<__nv_bfloat16*>(q_out_dev), static_cast<__nv_bfloat16*>(k_out_dev),
                                        position_ids_dev, inv_freq_dev, a);
  else
    mrope_kernel<<<grid, block, 0, s>>>(static_cast<const float*>(q_dev), static_cast<const float*>(k_dev),
                                        static_cast<float*>(q_out_dev), static_cast<float*>(k_out_dev), position_ids_dev,
                                        inv_freq_dev, a);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}
