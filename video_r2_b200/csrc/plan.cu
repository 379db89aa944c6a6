// Index plan construction + the bit-exact row gather.
//
// Closed forms (validated against HF get_window_index / rot_pos_emb by
// tests/test_oracle_indices.py through oracle/indices.py):
//   merged grid (lh, lw) = (h/2, w/2); windows are 4x4 merged tokens, visited
//   t -> window row a -> window col b -> i -> j, keeping 4a+i < lh and 4b+j < lw
//   (HF modeling_qwen2_5_vl.py:411-451 pads with -100 and filters; the padded,
//   empty windows vanish in unique_consecutive, :476).
// One thread per patch row of the WINDOW-ordered stream inverts that order in
// O(1), so no sort, no host loop and no .tolist().
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "plan.cuh"

namespace vr2 {

namespace {

struct ClipDesc {
  int64_t tok_base;    // first merged token of the clip (original order)
  int32_t t, lh, lw;   // temporal slices, merged rows, merged cols
  int32_t pad;
};

__global__ void plan_tables_kernel(const ClipDesc* __restrict__ clips, int n_clips, int64_t n_tokens, int ws,
                                   const float* __restrict__ inv_freq, int half_dim, int64_t* __restrict__ window_index,
                                   int64_t* __restrict__ reverse_index, int32_t* __restrict__ row_map,
                                   int32_t* __restrict__ out_map, int32_t* __restrict__ pos_hw,
                                   float* __restrict__ cos_tab, float* __restrict__ sin_tab,
                                   int32_t* __restrict__ pos_win) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // window-order patch row
  if (r >= n_tokens * 4) return;
  const int64_t p = r >> 2;
  const int m = static_cast<int>(r & 3);
  // clip lookup (clips sorted by tok_base)
  int lo = 0, hi = n_clips - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (clips[mid].tok_base <= p) lo = mid; else hi = mid - 1;
  }
  const ClipDesc c = clips[lo];
  const int64_t q = p - c.tok_base;
  const int64_t per_slice = static_cast<int64_t>(c.lh) * c.lw;
  const int64_t s = q / per_slice;
  int rem = static_cast<int>(q - s * per_slice);
  const int a = rem / (ws * c.lw);
  rem -= a * ws * c.lw;
  const int rows_a = min(ws, c.lh - a * ws);
  const int b = rem / (rows_a * ws);
  rem -= b * rows_a * ws;
  const int cols_b = min(ws, c.lw - b * ws);
  const int i = rem / cols_b;
  const int j = rem - i * cols_b;
  const int th = a * ws + i, tw = b * ws + j;  // merged-token coordinates
  const int64_t src_tok = c.tok_base + s * per_slice + static_cast<int64_t>(th) * c.lw + tw;
  const int64_t src_row = src_tok * 4 + m;
  const int hpos = 2 * th + (m >> 1), wpos = 2 * tw + (m & 1);

  if (m == 0) {
    window_index[p] = src_tok;
    reverse_index[src_tok] = p;
    out_map[p] = static_cast<int32_t>(src_tok);
  }
  row_map[src_row] = static_cast<int32_t>(r);
  pos_hw[src_row * 2 + 0] = hpos;
  pos_hw[src_row * 2 + 1] = wpos;
  pos_win[r] = hpos | (wpos << 16);
  const int quarter = half_dim / 2;
  for (int k = 0; k < quarter; ++k) {
    const float f = inv_freq[k];
    const float ah = static_cast<float>(hpos) * f;
    const float aw = static_cast<float>(wpos) * f;
    cos_tab[r * half_dim + k] = cosf(ah);
    sin_tab[r * half_dim + k] = sinf(ah);
    cos_tab[r * half_dim + quarter + k] = cosf(aw);
    sin_tab[r * half_dim + quarter + k] = sinf(aw);
  }
}

// The same angles per POSITION instead of per row: tab[p][k] = cos(p * f_k), tab[p][quarter + k] = sin(p * f_k) —
// bit-identical to the per-row tables (same expression), 160 bytes per position; the CTA-pair QKV GEMM keeps it in
// shared memory and looks rows up through pos_win.
__global__ void pos_table_kernel(const float* __restrict__ inv_freq, int quarter, int n_pos, float* __restrict__ tab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pos * quarter) return;
  const int p = i / quarter, k = i - p * quarter;
  const float a = static_cast<float>(p) * inv_freq[k];
  tab[p * 2 * quarter + k] = cosf(a);
  tab[p * 2 * quarter + quarter + k] = sinf(a);
}

// dst[i,:] = src[index[i],:], 16-byte vectors; one warp per row chunk keeps the
// loads and stores of a row contiguous.
__global__ void gather_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                   const int64_t* __restrict__ index, int64_t n_rows_out, int64_t vec_per_row) {
  const int64_t total = n_rows_out * vec_per_row;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = e / vec_per_row;
    const int64_t v = e - row * vec_per_row;
    dst[e] = __ldg(src + index[row] * vec_per_row + v);
  }
}

template <typename T>
T* carve(char*& cur, size_t count) {
  T* p = reinterpret_cast<T*>(cur);
  size_t bytes = (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
  cur += bytes;
  return p;
}

}  // namespace
}  // namespace vr2

using namespace vr2;

extern "C" int vr2_plan_create(const vr2_plan_config* cfg, const int64_t* grid, int32_t n_items, void* stream_v,
                               vr2_plan** out) {
  if (cfg == nullptr || grid == nullptr || out == nullptr || n_items <= 0)
    return fail(VR2_ERR_ARG, "vr2_plan_create: null argument or no grid rows");
  const int merge = cfg->spatial_merge_size;
  if (merge != 2) return fail(VR2_ERR_ARG, "vr2_plan_create: spatial_merge_size %d unsupported (2 only)", merge);
  const int ws = cfg->window_size / merge / cfg->patch_size;
  if (ws <= 0) return fail(VR2_ERR_ARG, "vr2_plan_create: window_size/merge/patch = %d", ws);
  if (cfg->head_dim % 4 != 0) return fail(VR2_ERR_ARG, "vr2_plan_create: head_dim %d must be a multiple of 4", cfg->head_dim);
  cudaStream_t stream = as_stream(stream_v);

  auto* plan = new vr2_plan();
  plan->merge = merge;
  plan->win_tokens = ws;
  plan->half_dim = cfg->head_dim / 2;
  plan->grid.assign(grid, grid + 3 * static_cast<size_t>(n_items));

  std::vector<ClipDesc> clips(n_items);
  std::vector<int32_t> cu_window{0}, cu_seq{0};
  int64_t tok = 0;
  for (int c = 0; c < n_items; ++c) {
    const int64_t t = grid[3 * c], h = grid[3 * c + 1], w = grid[3 * c + 2];
    if (t <= 0 || h <= 0 || w <= 0 || h % merge != 0 || w % merge != 0) {
      delete plan;
      return fail(VR2_ERR_ARG, "vr2_plan_create: grid_thw[%d] = (%lld,%lld,%lld): t,h,w must be positive and h,w divisible by %d",
                  c, (long long)t, (long long)h, (long long)w, merge);
    }
    const int lh = static_cast<int>(h / merge), lw = static_cast<int>(w / merge);
    clips[c] = ClipDesc{tok, static_cast<int32_t>(t), lh, lw, 0};
    const int nwh = (lh + ws - 1) / ws, nww = (lw + ws - 1) / ws;
    for (int64_t s = 0; s < t; ++s) {
      // full-attention segment = one temporal slice of h*w patches (HF :488-496)
      const int64_t seg0 = cu_seq.back();
      const int64_t L = h * w;
      cu_seq.push_back(static_cast<int32_t>(seg0 + L));
      for (int a = 0; a < nwh; ++a)
        for (int b = 0; b < nww; ++b) {
          const int rows = std::min(ws, lh - a * ws), cols = std::min(ws, lw - b * ws);
          const int len = rows * cols * merge * merge;  // patches in this window (<= 64)
          const int w0 = cu_window.back();
          cu_window.push_back(w0 + len);
        }
    }
    tok += t * lh * lw;
  }
  if (tok * merge * merge > INT32_MAX) {
    delete plan;
    return fail(VR2_ERR_ARG, "vr2_plan_create: %lld patches exceed int32 row indexing", (long long)(tok * 4));
  }
  plan->n_tokens = tok;
  plan->n_patches = tok * merge * merge;
  plan->n_windows = static_cast<int64_t>(cu_window.size()) - 1;
  plan->n_slices = static_cast<int64_t>(cu_seq.size()) - 1;

  std::vector<float> inv_freq(plan->half_dim / 2);
  for (int k = 0; k < plan->half_dim / 2; ++k)  // HF :117-130: 1 / theta^(2k / (head_dim/2))
    inv_freq[k] = 1.0f / powf(cfg->rope_theta, static_cast<float>(2 * k) / static_cast<float>(plan->half_dim));

  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    delete plan;
    return fail(VR2_ERR_CUDA, "vr2_plan_create: no CUDA device (%s); there is no CPU path", cudaGetErrorString(e));
  }
  plan->device = dev;

  const size_t N = plan->n_patches, T = plan->n_tokens, HDIM = plan->half_dim;
  auto rnd = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  // per-row segment bounds for the tcgen05 attention kernel (attention2.cu): expanded from the cu arrays
  std::vector<int2> seg_w(N), seg_f(N);
  for (size_t i = 0; i + 1 < cu_window.size(); ++i)
    for (int r = cu_window[i]; r < cu_window[i + 1]; ++r) seg_w[r] = make_int2(cu_window[i], cu_window[i + 1]);
  for (size_t i = 0; i + 1 < cu_seq.size(); ++i)
    for (int r = cu_seq[i]; r < cu_seq[i + 1]; ++r) seg_f[r] = make_int2(cu_seq[i], cu_seq[i + 1]);
  int n_pos = 0;
  for (int c = 0; c < n_items; ++c) n_pos = std::max<int>(n_pos, static_cast<int>(std::max(grid[3 * c + 1], grid[3 * c + 2])));
  plan->n_pos = n_pos;
  const size_t total = rnd(N * 4) + rnd(static_cast<size_t>(n_pos) * HDIM * 4) + rnd(N * 8) * 2 + rnd(T * 8) * 2 + rnd(N * 4) + rnd(T * 4) + rnd(N * 8) + rnd(N * HDIM * 4) * 2 +
                       rnd(cu_window.size() * 4) + rnd(cu_seq.size() * 4) +
                       rnd(clips.size() * sizeof(ClipDesc)) + rnd(inv_freq.size() * 4);
  e = cudaMalloc(&plan->blob, total);
  if (e != cudaSuccess) {
    delete plan;
    return fail(VR2_ERR_CUDA, "vr2_plan_create: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
  }
  char* cur = static_cast<char*>(plan->blob);
  plan->window_index = carve<int64_t>(cur, T);
  plan->reverse_index = carve<int64_t>(cur, T);
  plan->row_map = carve<int32_t>(cur, N);
  plan->out_map = carve<int32_t>(cur, T);
  plan->pos_hw = carve<int32_t>(cur, N * 2);
  plan->cos_tab = carve<float>(cur, N * HDIM);
  plan->sin_tab = carve<float>(cur, N * HDIM);
  plan->cu_window = carve<int32_t>(cur, cu_window.size());
  plan->cu_seqlens = carve<int32_t>(cur, cu_seq.size());
  plan->pos_win = carve<int32_t>(cur, N);
  plan->pos_tab = carve<float>(cur, static_cast<size_t>(n_pos) * HDIM);
  plan->seg_window = carve<int2>(cur, N);
  plan->seg_full = carve<int2>(cur, N);
  ClipDesc* d_clips = carve<ClipDesc>(cur, clips.size());
  float* d_inv = carve<float>(cur, inv_freq.size());

  auto up = [&](void* dst, const void* src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream); };
  e = up(plan->cu_window, cu_window.data(), cu_window.size() * 4);
  if (e == cudaSuccess) e = up(plan->cu_seqlens, cu_seq.data(), cu_seq.size() * 4);
  if (e == cudaSuccess) e = up(plan->seg_window, seg_w.data(), N * 8);
  if (e == cudaSuccess) e = up(plan->seg_full, seg_f.data(), N * 8);
  if (e == cudaSuccess) e = up(d_clips, clips.data(), clips.size() * sizeof(ClipDesc));
  if (e == cudaSuccess) e = up(d_inv, inv_freq.data(), inv_freq.size() * 4);
  if (e == cudaSuccess) {
    const int threads = 256;
    const int64_t blocks = (static_cast<int64_t>(N) + threads - 1) / threads;
    plan_tables_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
        d_clips, n_items, plan->n_tokens, ws, d_inv, plan->half_dim, plan->window_index, plan->reverse_index,
        plan->row_map, plan->out_map, plan->pos_hw, plan->cos_tab, plan->sin_tab, plan->pos_win);
    e = cudaGetLastError();
    if (e == cudaSuccess) {
      const int quarter = plan->half_dim / 2;
      pos_table_kernel<<<(n_pos * quarter + 255) / 256, 256, 0, stream>>>(d_inv, quarter, n_pos, plan->pos_tab);
      e = cudaGetLastError();
    }
  }
  // the host staging vectors die with this frame: make the copies land first
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    cudaFree(plan->blob);
    delete plan;
    return fail(VR2_ERR_CUDA, "vr2_plan_create: %s", cudaGetErrorString(e));
  }
  *out = plan;
  return VR2_OK;
}

extern "C" void vr2_plan_destroy(vr2_plan* plan) {
  if (plan == nullptr) return;
  if (plan->blob != nullptr) cudaFree(plan->blob);
  delete plan;
}

extern "C" int64_t vr2_plan_num_patches(const vr2_plan* p) { return p ? p->n_patches : -1; }
extern "C" int64_t vr2_plan_num_tokens(const vr2_plan* p) { return p ? p->n_tokens : -1; }
extern "C" int64_t vr2_plan_num_windows(const vr2_plan* p) { return p ? p->n_windows : -1; }
extern "C" int64_t vr2_plan_num_slices(const vr2_plan* p) { return p ? p->n_slices : -1; }

extern "C" int vr2_plan_export(const vr2_plan* p, int64_t* window_index, int64_t* reverse_index, int32_t* cu_window,
                               int32_t* cu_seqlens, int32_t* pos_hw, void* stream_v) {
  if (p == nullptr) return fail(VR2_ERR_ARG, "vr2_plan_export: null plan");
  cudaStream_t s = as_stream(stream_v);
  auto cp = [&](void* dst, const void* src, size_t bytes) {
    return dst == nullptr ? cudaSuccess : cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
  };
  VR2_CUDA(cp(window_index, p->window_index, p->n_tokens * 8));
  VR2_CUDA(cp(reverse_index, p->reverse_index, p->n_tokens * 8));
  VR2_CUDA(cp(cu_window, p->cu_window, (p->n_windows + 1) * 4));
  VR2_CUDA(cp(cu_seqlens, p->cu_seqlens, (p->n_slices + 1) * 4));
  VR2_CUDA(cp(pos_hw, p->pos_hw, p->n_patches * 8));
  return VR2_OK;
}

extern "C" int vr2_plan_export_rotary(const vr2_plan* p, float* cos_dev, float* sin_dev, void* stream_v) {
  if (p == nullptr) return fail(VR2_ERR_ARG, "vr2_plan_export_rotary: null plan");
  cudaStream_t s = as_stream(stream_v);
  const size_t bytes = static_cast<size_t>(p->n_patches) * p->half_dim * 4;
  if (cos_dev) VR2_CUDA(cudaMemcpyAsync(cos_dev, p->cos_tab, bytes, cudaMemcpyDeviceToDevice, s));
  if (sin_dev) VR2_CUDA(cudaMemcpyAsync(sin_dev, p->sin_tab, bytes, cudaMemcpyDeviceToDevice, s));
  return VR2_OK;
}

extern "C" int vr2_gather_rows(const void* src, void* dst, const int64_t* index, int64_t n_rows_out, int64_t n_rows_in,
                               int64_t row_bytes, void* stream_v) {
  if (n_rows_out == 0) return VR2_OK;
  if (src == nullptr || dst == nullptr || index == nullptr || n_rows_out < 0 || n_rows_in <= 0)
    return fail(VR2_ERR_ARG, "vr2_gather_rows: null pointer or negative size");
  if (row_bytes <= 0 || row_bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    return fail(VR2_ERR_ARG, "vr2_gather_rows: rows must be 16-byte multiples and 16-byte aligned (row_bytes=%lld)", (long long)row_bytes);
  const int64_t vec = row_bytes / 16;
  const int64_t total = n_rows_out * vec;
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride, a multiple of the SM count
  gather_rows_kernel<<<static_cast<unsigned>(blocks), threads, 0, as_stream(stream_v)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), index, n_rows_out, vec);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}
