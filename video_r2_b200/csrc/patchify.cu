// Frames -> the tower's input rows (the part after the resize): the rescale + normalise + patchify
// that Hugging Face's Qwen2VLVideoProcessor._preprocess does on the CPU
// (transformers video_processing_qwen2_vl.py: rescale_and_normalize -> pad T to a multiple of temporal_patch_size by
// repeating the last frame -> view [gt, Tp, C, gh/2, 2, 14, gw/2, 2, 14] -> permute(0,1,4,7,5,8,3,2,6,9) -> [N, 1176];
// Video-R2 reaches it from train/src/dataset and eval through the processor call), followed by the caller's
// `.type(self.visual.dtype)` (monkey_patch_forward.py:391).
//
//   out[row, ((c*Tp + tp)*P + ph)*P + pw] = (float(frame[min(gt*Tp + tp, T-1), c, (2*bh+mh)*P + ph, (2*bw+mw)*P + pw]) - mean[c]) / std[c]
//   row = ((gt*(gh/2) + bh)*(gw/2) + bw)*4 + mh*2 + mw
//
// with mean / std already fused with the rescale factor by the host exactly as HF does
// (image_processing_utils_fast.py _fuse_mean_std_and_rescale_factor).  fp32 subtract then IEEE divide, like
// torchvision's normalize (tensor.sub_(mean).div_(std)): bit-exact.  HBM-bound: uint8 frames in (1 B/pixel), bf16
// rows out (2 B/pixel); the upload shrinks from 4 (fp32) or 2 (bf16) bytes per pixel to 1.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vr2 {

namespace {

struct PatchifyArgs {
  int T, C, H, W;       // frames as given (before temporal padding)
  int P, Tp, merge;     // patch 14, temporal patch 2, merge 2
  int gt, gh, gw;       // grid
  float mean[4], std[4];
};

// One warp per output row (1176 elements for C=3, Tp=2, P=14): the row's coordinates are decomposed once, lanes walk
// the row two elements at a time, so stores are 128 contiguous bytes per instruction (bf16) and loads are the 14-pixel
// runs of the source patch rows.  P and Tp are compile-time for the Qwen2.5-VL geometry (divisions by constants).
template <typename In, typename Out, int P, int TP>
__global__ void patchify_kernel(const In* __restrict__ frames, Out* __restrict__ out, const PatchifyArgs a, int64_t rows) {
  const int p = P > 0 ? P : a.P, tpn = TP > 0 ? TP : a.Tp;
  const int pairs_per_row = a.C * tpn * p * p / 2;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int bwn = a.gw / a.merge, bhn = a.gh / a.merge;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    int64_t r = row;
    const int mw = static_cast<int>(r % a.merge); r /= a.merge;
    const int mh = static_cast<int>(r % a.merge); r /= a.merge;
    const int bw = static_cast<int>(r % bwn); r /= bwn;
    const int bh = static_cast<int>(r % bhn); r /= bhn;
    const int gt = static_cast<int>(r);
    const int y0 = (bh * a.merge + mh) * p, x0 = (bw * a.merge + mw) * p;
    for (int i = lane; i < pairs_per_row; i += 32) {
      int e = 2 * i;                         // even element of the row; P is even, so e + 1 shares (c, tp, ph)
      const int pw = e % p; e /= p;
      const int ph = e % p; e /= p;
      const int tp = e % tpn;
      const int c = e / tpn;
      const int t = min(gt * tpn + tp, a.T - 1);   // temporal padding repeats the last frame
      const In* src = frames + ((static_cast<int64_t>(t) * a.C + c) * a.H + (y0 + ph)) * a.W + x0 + pw;
      const float v0 = __fdiv_rn(__fsub_rn(static_cast<float>(src[0]), a.mean[c]), a.std[c]);
      const float v1 = __fdiv_rn(__fsub_rn(static_cast<float>(src[1]), a.mean[c]), a.std[c]);
      if constexpr (sizeof(Out) == 2) {
        reinterpret_cast<uint32_t*>(out)[row * pairs_per_row + i] = pack_bf16x2(v0, v1);
      } else {
        reinterpret_cast<float2*>(out)[row * pairs_per_row + i] = make_float2(v0, v1);
      }
    }
  }
}

template <typename In, typename Out>
void launch_patchify(const In* frames, Out* out, const PatchifyArgs& a, int64_t rows, cudaStream_t s) {
  int64_t blocks = (rows * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const unsigned g = static_cast<unsigned>(blocks);
  if (a.P == 14 && a.Tp == 2)
    patchify_kernel<In, Out, 14, 2><<<g, 256, 0, s>>>(frames, out, a, rows);
  else
    patchify_kernel<In, Out, 0, 0><<<g, 256, 0, s>>>(frames, out, a, rows);
}

}  // namespace
}  // namespace vr2

using namespace vr2;

extern "C" int vr2_patchify(const void* frames_dev, int32_t frames_dtype, int32_t T, int32_t C, int32_t H, int32_t W,
                            int32_t patch_size, int32_t temporal_patch_size, int32_t merge_size, const float* mean_host,
                            const float* std_host, void* out_dev, int32_t out_dtype, void* stream_v) {
  if (frames_dev == nullptr || out_dev == nullptr || mean_host == nullptr || std_host == nullptr)
    return fail(VR2_ERR_ARG, "vr2_patchify: null pointer");
  if (frames_dtype != VR2_U8 && frames_dtype != VR2_F32) return fail(VR2_ERR_ARG, "vr2_patchify: frames must be uint8 or fp32");
  if (out_dtype != VR2_F32 && out_dtype != VR2_BF16) return fail(VR2_ERR_ARG, "vr2_patchify: out must be fp32 or bf16");
  if (T <= 0 || C <= 0 || C > 4 || patch_size <= 0 || patch_size % 2 != 0 || temporal_patch_size <= 0 || merge_size <= 0)
    return fail(VR2_ERR_ARG, "vr2_patchify: bad T/C/patch/temporal patch/merge (%d,%d,%d,%d,%d)", T, C, patch_size,
                temporal_patch_size, merge_size);
  if (H <= 0 || W <= 0 || H % (patch_size * merge_size) != 0 || W % (patch_size * merge_size) != 0)
    return fail(VR2_ERR_ARG, "vr2_patchify: frame size %dx%d is not a multiple of patch*merge = %d", H, W, patch_size * merge_size);
  PatchifyArgs a;
  a.T = T; a.C = C; a.H = H; a.W = W;
  a.P = patch_size; a.Tp = temporal_patch_size; a.merge = merge_size;
  a.gt = (T + temporal_patch_size - 1) / temporal_patch_size;
  a.gh = H / patch_size; a.gw = W / patch_size;
  for (int c = 0; c < 4; ++c) { a.mean[c] = c < C ? mean_host[c] : 0.f; a.std[c] = c < C ? std_host[c] : 1.f; }
  const int64_t rows = static_cast<int64_t>(a.gt) * a.gh * a.gw;
  cudaStream_t s = as_stream(stream_v);
  if (frames_dtype == VR2_U8 && out_dtype == VR2_BF16)
    launch_patchify(static_cast<const uint8_t*>(frames_dev), static_cast<__nv_bfloat16*>(out_dev), a, rows, s);
  else if (frames_dtype == VR2_U8)
    launch_patchify(static_cast<const uint8_t*>(frames_dev), static_cast<float*>(out_dev), a, rows, s);
  else if (out_dtype == VR2_BF16)
    launch_patchify(static_cast<const float*>(frames_dev), static_cast<__nv_bfloat16*>(out_dev), a, rows, s);
  else
    launch_patchify(static_cast<const float*>(frames_dev), static_cast<float*>(out_dev), a, rows, s);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}
