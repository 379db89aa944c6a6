// get_rope_index with transformers-4.51.1 semantics (what Video-R2 pins,
// requirements.txt:3; call sites train/src/train/monkey_patch_forward.py:415-421
// and :583-589).  The restated algorithm is oracle/rope_index.py; the CUDA form:
//
//   kernel 1  one CTA per batch row: compact the unmasked tokens (block scan),
//             count the row's images / videos (vision_start followed by the
//             image / video token)
//   kernel 2  one CTA per batch row: exclusive prefix of those counts gives the
//             row's first image / video index (the reference threads
//             image_index / video_index through the batch loop); then the
//             row's vision items are walked in order — each step is a
//             block-wide "first image/video token at or after st" search
//             followed by a parallel fill of the text run and of the (t,h,w)
//             grid — and the trailing text is filled.
// position_ids is initialised to 1 (torch.ones), padding stays 1; the delta uses
// the PADDED row length S, as 4.51.1 does.  40 B of traffic per token.
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace vr2 {

namespace {

constexpr int kThreads = 512;

__device__ __forceinline__ bool mask_on(const void* mask, int dtype, int64_t idx) {
  if (mask == nullptr) return true;
  switch (dtype) {
    case VR2_I64: return static_cast<const int64_t*>(mask)[idx] == 1;
    case VR2_I32: return static_cast<const int32_t*>(mask)[idx] == 1;
    default: return static_cast<const uint8_t*>(mask)[idx] == 1;
  }
}
__device__ __forceinline__ int64_t mask_val(const void* mask, int dtype, int64_t idx) {
  switch (dtype) {
    case VR2_I64: return static_cast<const int64_t*>(mask)[idx];
    case VR2_I32: return static_cast<const int32_t*>(mask)[idx];
    default: return static_cast<const uint8_t*>(mask)[idx];
  }
}

// block-wide inclusive scan of one int per thread; returns the inclusive prefix, *total = block sum
__device__ int block_scan(int v, int* total, int* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (kThreads / 32) ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;  // inclusive over warps (kThreads/32 <= 32)
  }
  __syncthreads();
  const int base = warp > 0 ? warp_sums[warp - 1] : 0;
  *total = warp_sums[kThreads / 32 - 1];
  __syncthreads();
  return x + base;
}

// scratch layout: cidx [B,S] int32 (compact -> original position), counts [B,3] int32 (len, n_img, n_vid), then one
// int32 error word (bit 0: a row names more images than image_grid_thw has rows / no image grid at all, bit 1: the same
// for videos -- Hugging Face raises IndexError / TypeError there; vr2_rope_index_status reads it)
__global__ void __launch_bounds__(kThreads)
rope_compact_kernel(const int64_t* __restrict__ ids, const void* __restrict__ mask, int mask_dtype, int S,
                    int64_t vs_id, int64_t img_id, int64_t vid_id, int32_t* __restrict__ cidx,
                    int32_t* __restrict__ counts) {
  __shared__ int warp_sums[32];
  __shared__ int s_img, s_vid;
  const int b = blockIdx.x;
  const int64_t row = static_cast<int64_t>(b) * S;
  if (threadIdx.x == 0) { s_img = 0; s_vid = 0; }
  if (b == 0 && threadIdx.x == 0) counts[gridDim.x * 3] = 0;   // error word, set by rope_fill_kernel
  int running = 0;
  for (int base = 0; base < S; base += kThreads) {
    const int i = base + threadIdx.x;
    const int flag = (i < S && mask_on(mask, mask_dtype, row + i)) ? 1 : 0;
    int total;
    const int incl = block_scan(flag, &total, warp_sums);
    if (flag) cidx[row + running + incl - 1] = i;
    running += total;
  }
  __syncthreads();
  const int len = running;
  int n_img = 0, n_vid = 0;
  for (int j = threadIdx.x; j + 1 < len; j += kThreads) {
    if (ids[row + cidx[row + j]] == vs_id) {
      const int64_t nxt = ids[row + cidx[row + j + 1]];
      n_img += (nxt == img_id);
      n_vid += (nxt == vid_id);
    }
  }
  if (n_img) atomicAdd(&s_img, n_img);
  if (n_vid) atomicAdd(&s_vid, n_vid);
  __syncthreads();
  if (threadIdx.x == 0) {
    counts[b * 3 + 0] = len;
    counts[b * 3 + 1] = s_img;
    counts[b * 3 + 2] = s_vid;
  }
}

__global__ void __launch_bounds__(kThreads)
rope_fill_kernel(const int64_t* __restrict__ ids, int B, int S, const int64_t* __restrict__ image_thw, int n_img_total,
                 const int64_t* __restrict__ video_thw, int n_vid_total, const float* __restrict__ spg_arr,
                 vr2_rope_config cfg, const int32_t* __restrict__ cidx, const int32_t* __restrict__ counts,
                 int32_t* __restrict__ err, int64_t* __restrict__ pos, int64_t* __restrict__ deltas) {
  __shared__ int s_ed_img, s_ed_vid;
  const int b = blockIdx.x;
  const int64_t row = static_cast<int64_t>(b) * S;
  const int64_t plane = static_cast<int64_t>(B) * S;
  const int32_t* cid = cidx + row;
  // torch.ones(3, B, S): padding positions keep 1
  for (int i = threadIdx.x; i < S; i += kThreads) {
    pos[row + i] = 1;
    pos[plane + row + i] = 1;
    pos[2 * plane + row + i] = 1;
  }
  int img_i = 0, vid_i = 0;
  for (int r = 0; r < b; ++r) {
    img_i += counts[r * 3 + 1];
    vid_i += counts[r * 3 + 2];
  }
  const int len = counts[b * 3 + 0];
  int remain_img = counts[b * 3 + 1], remain_vid = counts[b * 3 + 2];
  const int n_items = remain_img + remain_vid;
  __syncthreads();  // the fill above and the writes below touch the same addresses

  int st = 0;
  int64_t next_pos = 0;   // llm_pos_ids_list[-1].max() + 1, or 0 while the list is empty
  int64_t max_pos = -1;
  for (int it = 0; it < n_items; ++it) {
    if (threadIdx.x == 0) { s_ed_img = len + 1; s_ed_vid = len + 1; }
    __syncthreads();
    // input_tokens.index(image_token_id, st) / .index(video_token_id, st): each thread's strided scan visits j
    // in increasing order, so its first hit is its minimum; the block minimum is the list's first occurrence.
    int ei = len + 1, ev = len + 1;
    for (int j = st + threadIdx.x; j < len; j += kThreads) {
      const int64_t tok = ids[row + cid[j]];
      if (ei > len && tok == cfg.image_token_id) ei = j;
      if (ev > len && tok == cfg.video_token_id) ev = j;
      if (ei <= len && ev <= len) break;
    }
    if (remain_img > 0 && ei <= len) atomicMin(&s_ed_img, ei);
    if (remain_vid > 0 && ev <= len) atomicMin(&s_ed_vid, ev);
    __syncthreads();
    const int ed_img = s_ed_img, ed_vid = s_ed_vid;
    __syncthreads();
    if (ed_img > len && ed_vid > len) break;  // malformed row (HF would raise): stop placing vision items
    int64_t t, h, w;
    float spg;
    int ed;
    // A row that names an item the grids do not describe (grid pointer null / empty, or more items than grid rows):
    // HF raises; here the item is not placed (no read through a null or out-of-range pointer), the rest of the row
    // is numbered as text and the error word is set.  All operands are block-uniform, so the break is too.
    if (ed_img < ed_vid ? (image_thw == nullptr || img_i >= n_img_total) : (video_thw == nullptr || vid_i >= n_vid_total)) {
      if (threadIdx.x == 0) atomicOr(err, ed_img < ed_vid ? 1 : 2);
      break;
    }
    if (ed_img < ed_vid) {
      const int64_t* g = image_thw + 3 * static_cast<int64_t>(img_i);
      t = g[0]; h = g[1]; w = g[2];
      spg = 0.f;
      ++img_i; --remain_img; ed = ed_img;
    } else {
      const int vi = vid_i;
      const int64_t* g = video_thw + 3 * static_cast<int64_t>(vi);
      t = g[0]; h = g[1]; w = g[2];
      spg = spg_arr != nullptr ? spg_arr[vi] : 1.0f;
      ++vid_i; --remain_vid; ed = ed_vid;
    }
    const int gt = static_cast<int>(t);
    const int gh = static_cast<int>(h / cfg.spatial_merge_size), gw = static_cast<int>(w / cfg.spatial_merge_size);
    const int text_len = ed - st;
    const int64_t st_idx = next_pos;
    for (int j = st + threadIdx.x; j < ed; j += kThreads) {
      const int64_t v = st_idx + (j - st);
      const int o = cid[j];
      pos[row + o] = v;
      pos[plane + row + o] = v;
      pos[2 * plane + row + o] = v;
    }
    if (text_len > 0) max_pos = max(max_pos, st_idx + text_len - 1);
    const int64_t off = st_idx + text_len;
    const int64_t n_vis = static_cast<int64_t>(gt) * gh * gw;
    // temporal index of slice k; must reproduce torch's arithmetic exactly
    auto t_index = [&](int k) -> int64_t {
      if (cfg.spg_int_cast) {
        const int64_t si = static_cast<int64_t>(spg);  // torch.as_tensor(spg, dtype=int64): trunc toward zero
        const float tps = cfg.tokens_per_second;
        if (tps == truncf(tps)) return static_cast<int64_t>(k) * si * static_cast<int64_t>(tps);
        return static_cast<int64_t>(__fmul_rn(static_cast<float>(static_cast<int64_t>(k) * si), tps));
      }
      // int64 * float32 -> float32; * tokens_per_second -> float32; .long() truncates
      const float a = __fmul_rn(static_cast<float>(k), spg);
      return static_cast<int64_t>(__fmul_rn(a, cfg.tokens_per_second));
    };
    for (int64_t q = threadIdx.x; q < n_vis; q += kThreads) {
      const int64_t j = ed + q;
      if (j >= len) break;
      const int k = static_cast<int>(q / (static_cast<int64_t>(gh) * gw));
      const int rem = static_cast<int>(q - static_cast<int64_t>(k) * gh * gw);
      const int o = cid[j];
      pos[row + o] = t_index(k) + off;
      pos[plane + row + o] = rem / gw + off;
      pos[2 * plane + row + o] = rem % gw + off;
    }
    if (n_vis > 0) {
      int64_t vmax = max(static_cast<int64_t>(gh - 1), static_cast<int64_t>(gw - 1));
      // t_index is monotone in k for spg >= 0; take both ends so a negative spg cannot hide the max
      vmax = max(vmax, max(t_index(gt - 1), t_index(0)));
      max_pos = max(max_pos, vmax + off);
      next_pos = vmax + off + 1;
    } else if (text_len > 0) {
      next_pos = st_idx + text_len;
    }
    st = ed + static_cast<int>(n_vis);
  }
  if (st < len) {
    const int64_t st_idx = next_pos;
    for (int j = st + threadIdx.x; j < len; j += kThreads) {
      const int64_t v = st_idx + (j - st);
      const int o = cid[j];
      pos[row + o] = v;
      pos[plane + row + o] = v;
      pos[2 * plane + row + o] = v;
    }
    max_pos = max(max_pos, st_idx + (len - st) - 1);
  }
  if (threadIdx.x == 0) deltas[b] = max_pos + 1 - S;
}

// Text-only branch with an attention mask: cumsum(mask) - 1, masked_fill(mask == 0, 1).
__global__ void __launch_bounds__(kThreads)
rope_text_mask_kernel(const void* __restrict__ mask, int mask_dtype, int B, int S, int64_t* __restrict__ pos,
                      int64_t* __restrict__ deltas) {
  __shared__ int warp_sums[32];
  __shared__ long long s_max;
  const int b = blockIdx.x;
  const int64_t row = static_cast<int64_t>(b) * S;
  const int64_t plane = static_cast<int64_t>(B) * S;
  if (threadIdx.x == 0) s_max = INT64_MIN;
  __syncthreads();
  int64_t running = 0;
  long long local_max = INT64_MIN;
  for (int base = 0; base < S; base += kThreads) {
    const int i = base + threadIdx.x;
    const int64_t mv = i < S ? mask_val(mask, mask_dtype, row + i) : 0;
    int total;
    const int incl = block_scan(static_cast<int>(mv), &total, warp_sums);
    if (i < S) {
      const int64_t v = mv == 0 ? 1 : running + incl - 1;
      pos[row + i] = v;
      pos[plane + row + i] = v;
      pos[2 * plane + row + i] = v;
      local_max = max(local_max, static_cast<long long>(v));
    }
    running += total;
  }
  atomicMax(&s_max, local_max);
  __syncthreads();
  if (threadIdx.x == 0) deltas[b] = static_cast<int64_t>(s_max) + 1 - S;
}

__global__ void rope_text_plain_kernel(int B, int S, int64_t* __restrict__ pos, int64_t* __restrict__ deltas) {
  const int64_t total = static_cast<int64_t>(B) * S;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t v = e % S;
    pos[e] = v;
    pos[total + e] = v;
    pos[2 * total + e] = v;
    if (e < B) deltas[e] = 0;
  }
}

}  // namespace
}  // namespace vr2

using namespace vr2;

extern "C" size_t vr2_rope_scratch_bytes(int32_t B, int32_t S) {
  return (static_cast<size_t>(B) * S + static_cast<size_t>(B) * 3 + 8) * sizeof(int32_t);
}

extern "C" int vr2_rope_index(const vr2_rope_config* cfg, const int64_t* ids, const void* mask, int32_t mask_dtype,
                              int32_t B, int32_t S, const int64_t* image_thw, int32_t n_img, const int64_t* video_thw,
                              int32_t n_vid, const float* spg, void* scratch, int64_t* pos, int64_t* deltas,
                              void* stream_v) {
  if (cfg == nullptr || pos == nullptr || deltas == nullptr) return fail(VR2_ERR_ARG, "vr2_rope_index: null argument");
  if (B < 0 || S < 0) return fail(VR2_ERR_ARG, "vr2_rope_index: negative shape");
  if (B == 0 || S == 0) return VR2_OK;
  if (mask != nullptr && mask_dtype != VR2_I64 && mask_dtype != VR2_I32 && mask_dtype != VR2_U8)
    return fail(VR2_ERR_ARG, "vr2_rope_index: attention_mask dtype %d unsupported (int64/int32/uint8/bool)", mask_dtype);
  cudaStream_t stream = as_stream(stream_v);
  const bool have_grids = (image_thw != nullptr && n_img > 0) || (video_thw != nullptr && n_vid > 0);
  // HF branches on `image_grid_thw is not None or video_grid_thw is not None`; an empty grid tensor behaves like
  // a row without vision items, which the vision path below handles too.
  const bool vision_branch = ids != nullptr && (image_thw != nullptr || video_thw != nullptr);
  (void)have_grids;
  if (!vision_branch) {
    if (mask != nullptr) {
      rope_text_mask_kernel<<<B, kThreads, 0, stream>>>(mask, mask_dtype, B, S, pos, deltas);
    } else {
      int64_t total = static_cast<int64_t>(B) * S;
      int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
      rope_text_plain_kernel<<<blocks, 256, 0, stream>>>(B, S, pos, deltas);
    }
    VR2_CUDA(cudaGetLastError());
    return VR2_OK;
  }
  if (scratch == nullptr) return fail(VR2_ERR_ARG, "vr2_rope_index: scratch buffer required");
  if (cfg->spatial_merge_size <= 0) return fail(VR2_ERR_ARG, "vr2_rope_index: spatial_merge_size must be positive");
  int32_t* cidx = static_cast<int32_t*>(scratch);
  int32_t* counts = cidx + static_cast<size_t>(B) * S;
  rope_compact_kernel<<<B, kThreads, 0, stream>>>(ids, mask, mask_dtype, S, cfg->vision_start_token_id,
                                                   cfg->image_token_id, cfg->video_token_id, cidx, counts);
  VR2_CUDA(cudaGetLastError());
  rope_fill_kernel<<<B, kThreads, 0, stream>>>(ids, B, S, image_thw, n_img, video_thw, n_vid, spg, *cfg, cidx, counts,
                                                counts + static_cast<size_t>(B) * 3, pos, deltas);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}

extern "C" int vr2_rope_index_status(const void* scratch, int32_t B, int32_t S, int32_t* status_host, void* stream_v) {
  if (scratch == nullptr || status_host == nullptr || B <= 0 || S <= 0)
    return fail(VR2_ERR_ARG, "vr2_rope_index_status: null argument / empty shape");
  cudaStream_t stream = as_stream(stream_v);
  const int32_t* word = static_cast<const int32_t*>(scratch) + static_cast<size_t>(B) * S + static_cast<size_t>(B) * 3;
  VR2_CUDA(cudaMemcpyAsync(status_host, word, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  VR2_CUDA(cudaStreamSynchronize(stream));
  return VR2_OK;
}
