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
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
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

}  // namespace
}  // namespace vr2

using namespace vr2;

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
