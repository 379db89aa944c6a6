// Per-token log-probabilities of the LM head without materialising the logits:
//     logits = hidden @ W^T;  logp[m] = (logits[m] / T).log_softmax(-1)[target[m]]
// which Video-R2's GRPO computes three times per step over a 152 064-word vocabulary
// (train/src/trainer/grpo_trainer.py:723-741: model(...).logits -> logits / temperature -> selective_log_softmax),
// writing and re-reading a [B, L, V] bf16 tensor (2.5 GB for 8 192 tokens) each time.
// Here the CTA-pair tcgen05 GEMM (gemm2.cu, EPI_LOGPROB) reduces every 256-column block of the fp32 accumulator to a
// (max, sum-exp) pair per row and picks out the target's logit; this file combines the ceil(V/256) partials per row.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>

#include "common.cuh"
#include "gemm.cuh"

namespace vr2 {

namespace {

// one thread per token row; partials are part-major [parts][M]: coalesced across rows
__global__ void logprob_reduce_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                      const float* __restrict__ tgt_logit, float* __restrict__ out, int64_t M, int parts) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float mx = -INFINITY;
  for (int q = 0; q < parts; ++q) mx = fmaxf(mx, part_max[static_cast<int64_t>(q) * M + r]);
  float sum = 0.f;
  for (int q = 0; q < parts; ++q) {
    const float m = part_max[static_cast<int64_t>(q) * M + r];
    sum += part_sum[static_cast<int64_t>(q) * M + r] * (m == -INFINITY ? 0.f : expf(m - mx));
  }
  out[r] = tgt_logit[r] - (mx + logf(sum));
}

}  // namespace
}  // namespace vr2

using namespace vr2;

extern "C" size_t vr2_lm_logprob_workspace_bytes(int64_t M, int32_t V) {
  if (M <= 0 || V <= 0) return 0;
  const size_t parts = (static_cast<size_t>(V) + 255) / 256;
  return (2 * parts + 1) * static_cast<size_t>(M) * 4 + 256;
}

extern "C" int vr2_lm_logprob(const void* hidden_dev, const void* weight_dev, const int64_t* target_dev, int64_t M,
                              int32_t V, int32_t K, float temperature, float* out_logprob_dev, void* workspace_dev,
                              size_t workspace_bytes, void* stream_v) {
  if (hidden_dev == nullptr || weight_dev == nullptr || target_dev == nullptr || out_logprob_dev == nullptr ||
      workspace_dev == nullptr)
    return fail(VR2_ERR_ARG, "vr2_lm_logprob: null pointer");
  if (M <= 0 || M > INT32_MAX || V <= 0 || K <= 0 || K % 8 != 0)
    return fail(VR2_ERR_ARG, "vr2_lm_logprob: bad shape M=%lld V=%d K=%d (K must be a multiple of 8)", (long long)M, V, K);
  if (!(temperature > 0.f)) return fail(VR2_ERR_ARG, "vr2_lm_logprob: temperature must be positive");
  if (workspace_bytes < vr2_lm_logprob_workspace_bytes(M, V))
    return fail(VR2_ERR_ARG, "vr2_lm_logprob: workspace %zu < required %zu bytes", workspace_bytes, vr2_lm_logprob_workspace_bytes(M, V));
  if ((reinterpret_cast<uintptr_t>(workspace_dev) & 15) != 0) return fail(VR2_ERR_ARG, "vr2_lm_logprob: workspace must be 16-byte aligned");
  int dev = 0, sms = 148;
  VR2_CUDA(cudaGetDevice(&dev));
  VR2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaStream_t s = as_stream(stream_v);
  const int parts = (V + 255) / 256;
  float* part_max = static_cast<float*>(workspace_dev);
  float* part_sum = part_max + static_cast<size_t>(parts) * M;
  float* tgt = part_sum + static_cast<size_t>(parts) * M;
  VR2_CUDA(cudaMemsetAsync(tgt, 0xFF, static_cast<size_t>(M) * 4, s));   // NaN unless the target is a valid column
  GemmArgs a{};
  a.A = static_cast<const __nv_bfloat16*>(hidden_dev); a.lda = K;
  a.W = static_cast<const __nv_bfloat16*>(weight_dev); a.ldw = K;
  a.M = static_cast<int>(M); a.N = V; a.K = K;
  a.epilogue = EPI_LOGPROB;
  a.target = target_dev; a.inv_temp = 1.0f / temperature;
  a.part_max = part_max; a.part_sum = part_sum; a.tgt_logit = tgt;
  const char* msg = "";
  cudaError_t e = gemm_launch(a, sms, s, &msg);
  if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? VR2_ERR_ARG : VR2_ERR_CUDA, "vr2_lm_logprob: %s %s", cudaGetErrorString(e), msg);
  logprob_reduce_kernel<<<static_cast<unsigned>((M + 255) / 256), 256, 0, s>>>(part_max, part_sum, tgt, out_logprob_dev, M, parts);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}
