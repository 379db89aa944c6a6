// Frame resize on the GPU (the part before patchify): what Video-R2's fetch_video does on the host
// cores (train/qwen_vl_utils/vision_process.py:802-807, same code under demo/ and eval/):
//
//     video = transforms.functional.resize(video_u8[T,3,H,W], [h', w'], BICUBIC, antialias=True).float()
//
// i.e. torchvision's tensor path: uint8 -> float32, ATen's separable anti-aliased bicubic (width first, then height;
// Keys kernel a = -0.5; support 2*max(scale,1); weights normalised per output index; float32 throughout), clamp to
// [0,255], round half to even, back to uint8.  oracle/resize.py states the arithmetic (including the order of the
// multiply-adds the CPU library executes) and is pinned to torch bit for bit; this file follows it with explicit
// round-to-nearest intrinsics so that the bytes are identical, not merely close.
//
// Two kernels per call:
//   axis_weights_kernel  one thread per output column / row: first tap, tap count, normalised weights (tiny)
//   resize_kernel        one CTA per (frame, channel, 32x64 output tile): the uint8 source window is staged in shared
//                        memory once, the width pass writes a float32 strip [source rows][64] to shared memory, the
//                        height pass reads it and stores the tile.  The float32 intermediate of the two-pass form
//                        never reaches HBM: traffic is the source bytes (+ the halo rows shared by vertically
//                        adjacent tiles, L2 hits) and the output bytes.
// HBM-bound byte work (algorithmic bytes = T*C*(H*W + h'*w') for uint8 out).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace vr2 {

namespace {

struct AxisTable {
  const int* first;     // [out]
  const int* count;     // [out]
  const float* weight;  // [out][stride]
  int stride;           // max taps
};

__device__ __forceinline__ float keys_cubic(float x) {
  x = fabsf(x);
  if (x < 1.f) return __fmaf_rn(__fmul_rn(__fmaf_rn(1.5f, x, -2.5f), x), x, 1.f);
  if (x < 2.f) return __fmaf_rn(__fmaf_rn(__fmaf_rn(-0.5f, x, 2.5f), x, -4.f), x, 2.f);
  return 0.f;
}

__global__ void axis_weights_kernel(int in_size, int out_size, int max_taps, int* __restrict__ first, int* __restrict__ count,
                                    float* __restrict__ weight) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_size) return;
  const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
  const float support = scale >= 1.f ? __fmul_rn(2.f, scale) : 2.f;
  const float invscale = scale >= 1.f ? static_cast<float>(1.0 / static_cast<double>(scale)) : 1.f;
  const float center = static_cast<float>(__dmul_rn(static_cast<double>(scale), static_cast<double>(i) + 0.5));
  long long lo = static_cast<long long>(__dadd_rn(static_cast<double>(__fsub_rn(center, support)), 0.5));
  lo = lo < 0 ? 0 : lo;
  long long hi = static_cast<long long>(__dadd_rn(static_cast<double>(__fadd_rn(center, support)), 0.5));
  hi = hi > in_size ? in_size : hi;
  long long n = hi - lo;
  n = n < 0 ? 0 : (n > max_taps ? max_taps : n);
  float* w = weight + static_cast<size_t>(i) * max_taps;
  float total = 0.f;
  for (int j = 0; j < static_cast<int>(n); ++j) {
    const float d = __fsub_rn(static_cast<float>(j + lo), center);
    const float arg = static_cast<float>(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), static_cast<double>(invscale)));
    const float v = keys_cubic(arg);
    w[j] = v;
    total = __fadd_rn(total, v);
  }
  if (total != 0.f)
    for (int j = 0; j < static_cast<int>(n); ++j) w[j] = __fdiv_rn(w[j], total);
  for (int j = static_cast<int>(n); j < max_taps; ++j) w[j] = 0.f;
  first[i] = static_cast<int>(lo);
  count[i] = static_cast<int>(n);
}

struct ResizeArgs {
  const uint8_t* src;
  const uint8_t* src_end;    // one past the last byte of the frames view
  long long st, sc, sh, sw;  // source strides in bytes
  int C, H, W, oh, ow;
  AxisTable tx, ty;
  int tile_x, tile_y;        // output tile
  int cap_cols, cap_rows;    // capacity of the staged source window
};

// acc = x0*w0; then 4*floor((n-1)/4) taps as multiply, add; the last (n-1)%4 as fused multiply-adds (oracle/resize.py).
// R independent sums share each weight load (R source rows in the width pass, R adjacent columns in the height pass).
template <int R, typename X, typename Wt>
__device__ __forceinline__ void tap_sums(float (&acc)[R], X x, Wt w, int n) {
  if (n <= 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    return;
  }
  {
    const float w0 = w(0);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = __fmul_rn(x(r, 0), w0);
  }
  const int plain = ((n - 1) >> 2) << 2;
  int j = 1;
  for (; j <= plain; j += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float wj = w(j + u);
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(x(r, j + u), wj));
    }
  }
  for (; j < n; ++j) {
    const float wj = w(j);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = __fmaf_rn(x(r, j), wj, acc[r]);
  }
}

template <typename Out>
__device__ __forceinline__ Out finish(float v);
template <>
__device__ __forceinline__ uint8_t finish<uint8_t>(float v) {
  return static_cast<uint8_t>(__float2int_rn(fminf(fmaxf(v, 0.f), 255.f)));
}
template <>
__device__ __forceinline__ float finish<float>(float v) {
  return rintf(fminf(fmaxf(v, 0.f), 255.f));
}

constexpr int kThreads = 256;
constexpr int kRowBlock = 4;  // source rows per thread in the width pass
constexpr int kColBlock = 4;  // output columns per thread in the height pass

template <typename Out>
__global__ void __launch_bounds__(kThreads) resize_kernel(const ResizeArgs a, Out* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int ox0 = blockIdx.x * a.tile_x, oy0 = blockIdx.y * a.tile_y;
  const int nx = min(a.tile_x, a.ow - ox0), ny = min(a.tile_y, a.oh - oy0);
  const int plane = blockIdx.z;  // t*C + c
  const int t = plane / a.C, c = plane - t * a.C;

  // shared layout: wx [taps_x][tile_x] | wy [tile_y][taps_y] | mid [cap_rows + pad][tile_x] | win [cap_rows + pad][cap_cols]
  //                | fx, cx [tile_x] | fy, cy [tile_y]
  float* wx = reinterpret_cast<float*>(smem);
  float* wy = wx + a.tx.stride * a.tile_x;
  float* mid = wy + a.ty.stride * a.tile_y;
  float* win = mid + static_cast<size_t>(a.cap_rows + kRowBlock) * a.tile_x;
  int* fx = reinterpret_cast<int*>(win + static_cast<size_t>(a.cap_rows + kRowBlock) * a.cap_cols);
  int* cx = fx + a.tile_x;
  int* fy = cx + a.tile_x;
  int* cy = fy + a.tile_y;

  for (int i = tid; i < a.tile_x; i += kThreads) {
    const bool in = i < nx;                       // columns past the frame edge: zero taps
    fx[i] = in ? a.tx.first[ox0 + i] : 0;
    cx[i] = in ? a.tx.count[ox0 + i] : 0;
  }
  for (int i = tid; i < ny; i += kThreads) { fy[i] = a.ty.first[oy0 + i]; cy[i] = a.ty.count[oy0 + i]; }
  for (int j = warp; j < a.tx.stride; j += kThreads / 32)  // transposed: in the width pass lanes walk x
    for (int x = lane; x < nx; x += 32) wx[j * a.tile_x + x] = a.tx.weight[static_cast<size_t>(ox0 + x) * a.tx.stride + j];
  for (int i = tid; i < a.ty.stride * ny; i += kThreads) wy[i] = a.ty.weight[static_cast<size_t>(oy0) * a.ty.stride + i];
  __syncthreads();
  // first taps and (first + count) are non-decreasing in the output index
  const int col0 = fx[0], col1 = fx[nx - 1] + cx[nx - 1];
  const int row0 = fy[0], row1 = fy[ny - 1] + cy[ny - 1];
  const int cols = col1 - col0, rows = row1 - row0;
  if (cols > a.cap_cols || rows > a.cap_rows) __trap();  // the host's bound on the window was wrong: fail loudly

  // stage the source window as float32 (one conversion per source byte instead of one per tap); one warp per row
  const uint8_t* base = a.src + t * a.st + c * a.sc;
  const int pitch = a.cap_cols;
  const int rows_pad = (rows + kRowBlock - 1) / kRowBlock * kRowBlock;
  for (int r = warp; r < rows_pad; r += kThreads / 32) {
    float* dst = win + r * pitch;
    if (r >= rows) {                               // padding rows of the last row block: defined values, never stored
      for (int x = lane; x < cols; x += 32) dst[x] = 0.f;
      continue;
    }
    const uint8_t* src = base + (row0 + r) * a.sh + col0 * a.sw;
    if (a.sw == 1) {
      // contiguous pixels: aligned 4-byte loads; a word that would reach outside the frames view is read bytewise
      const uintptr_t p = reinterpret_cast<uintptr_t>(src);
      const int lead = static_cast<int>(p & 3);
      const uint8_t* word0 = src - lead;
      const int nwords = (lead + cols + 3) >> 2;
      for (int wi = lane; wi < nwords; wi += 32) {
        const uint8_t* wp = word0 + wi * 4;
        const int x = wi * 4 - lead;
        if (wp >= a.src && wp + 4 <= a.src_end) {
          const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(wp));
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (x + b >= 0 && x + b < cols) dst[x + b] = static_cast<float>((v >> (8 * b)) & 0xffu);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (x + b >= 0 && x + b < cols) dst[x + b] = static_cast<float>(__ldg(wp + b));
        }
      }
    } else {
      for (int x = lane; x < cols; x += 32) dst[x] = static_cast<float>(__ldg(src + x * a.sw));
    }
  }
  __syncthreads();

  // width pass: thread = one output column x, kRowBlock source rows at a time -> mid[r][x]
  {
    const int x = tid & (a.tile_x - 1);
    const int rgroups = kThreads / a.tile_x;       // tile_x is a power of two <= 64
    const int n = cx[x];
    const float* px = win + (fx[x] - col0);
    const float* w = wx + x;
    const int wstride = a.tile_x;
    for (int r0 = (tid / a.tile_x) * kRowBlock; r0 < rows; r0 += rgroups * kRowBlock) {
      float acc[kRowBlock];
      const float* p = px + r0 * pitch;
      tap_sums<kRowBlock>(acc, [&](int r, int j) { return p[r * pitch + j]; }, [&](int j) { return w[j * wstride]; }, n);
#pragma unroll
      for (int r = 0; r < kRowBlock; ++r) mid[(r0 + r) * a.tile_x + x] = acc[r];
    }
  }
  __syncthreads();

  // height pass + clamp + round + store: thread = kColBlock columns (x, x + q, x + 2q, ..., q = tile_x / kColBlock) of one
  // output row; they share the weights and the row bookkeeping
  {
    Out* dst = out + (static_cast<size_t>(plane) * a.oh + oy0) * a.ow + ox0;
    if (a.tile_x >= kColBlock) {
      const int q = a.tile_x / kColBlock;
      const int xq = tid & (q - 1);
      if (xq < nx) {
        for (int y = tid / q; y < ny; y += kThreads / q) {
          const float* col = mid + (fy[y] - row0) * a.tile_x + xq;
          const float* w = wy + y * a.ty.stride;
          const int stride = a.tile_x;
          float acc[kColBlock];
          tap_sums<kColBlock>(acc, [&](int r, int j) { return col[j * stride + r * q]; }, [&](int j) { return w[j]; }, cy[y]);
          Out* row = dst + static_cast<size_t>(y) * a.ow + xq;
#pragma unroll
          for (int r = 0; r < kColBlock; ++r)
            if (xq + r * q < nx) row[r * q] = finish<Out>(acc[r]);
        }
      }
    } else {
      for (int i = tid; i < ny * a.tile_x; i += kThreads) {
        const int y = i / a.tile_x, x = i - y * a.tile_x;
        if (x >= nx) continue;
        const float* col = mid + (fy[y] - row0) * a.tile_x + x;
        const float* w = wy + y * a.ty.stride;
        const int stride = a.tile_x;
        float acc[1];
        tap_sums<1>(acc, [&](int, int j) { return col[j * stride]; }, [&](int j) { return w[j]; }, cy[y]);
        dst[static_cast<size_t>(y) * a.ow + x] = finish<Out>(acc[0]);
      }
    }
  }
}

int max_taps(int in_size, int out_size) {
  const float scale = static_cast<float>(in_size) / static_cast<float>(out_size);
  const float support = scale >= 1.f ? 2.f * scale : 2.f;
  return static_cast<int>(std::ceil(support)) * 2 + 1;
}

// upper bound on (first(last) + count(last) - first(first)) over a run of `tile` consecutive outputs
int window_cap(int in_size, int out_size, int tile) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double support = scale >= 1.0 ? 2.0 * scale : 2.0;
  const int cap = static_cast<int>(std::ceil((tile - 1) * scale + 2.0 * support)) + 3;
  return std::min(cap, in_size);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TableLayout {
  int taps_x, taps_y;
  size_t off_fx, off_cx, off_wx, off_fy, off_cy, off_wy, bytes;
};

TableLayout table_layout(int H, int W, int oh, int ow) {
  TableLayout l;
  l.taps_x = max_taps(W, ow);
  l.taps_y = max_taps(H, oh);
  size_t o = 0;
  l.off_fx = o; o = align_up(o + sizeof(int) * ow, 256);
  l.off_cx = o; o = align_up(o + sizeof(int) * ow, 256);
  l.off_wx = o; o = align_up(o + sizeof(float) * static_cast<size_t>(ow) * l.taps_x, 256);
  l.off_fy = o; o = align_up(o + sizeof(int) * oh, 256);
  l.off_cy = o; o = align_up(o + sizeof(int) * oh, 256);
  l.off_wy = o; o = align_up(o + sizeof(float) * static_cast<size_t>(oh) * l.taps_y, 256);
  l.bytes = o;
  return l;
}

size_t tile_smem(const TableLayout& l, int tile_x, int tile_y, int cap_rows, int cap_cols) {
  return sizeof(float) * (static_cast<size_t>(l.taps_x) * tile_x + static_cast<size_t>(l.taps_y) * tile_y +
                          static_cast<size_t>(cap_rows + kRowBlock) * (tile_x + cap_cols)) +
         sizeof(int) * 2 * (tile_x + tile_y) + 16;
}

}  // namespace

}  // namespace vr2

using namespace vr2;

extern "C" int vr2_resize_workspace_bytes(int32_t H, int32_t W, int32_t out_h, int32_t out_w, size_t* bytes) {
  if (bytes == nullptr) return fail(VR2_ERR_ARG, "vr2_resize_workspace_bytes: null pointer");
  if (H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0) return fail(VR2_ERR_ARG, "vr2_resize_workspace_bytes: sizes must be positive");
  *bytes = table_layout(H, W, out_h, out_w).bytes;
  return VR2_OK;
}

extern "C" int vr2_resize_frames(const void* frames_dev, int32_t T, int32_t C, int32_t H, int32_t W, int64_t stride_t,
                                 int64_t stride_c, int64_t stride_h, int64_t stride_w, int32_t out_h, int32_t out_w,
                                 void* out_dev, int32_t out_dtype, void* workspace_dev, void* stream_v) {
  if (frames_dev == nullptr || out_dev == nullptr || workspace_dev == nullptr)
    return fail(VR2_ERR_ARG, "vr2_resize_frames: null pointer");
  if (out_dtype != VR2_U8 && out_dtype != VR2_F32) return fail(VR2_ERR_ARG, "vr2_resize_frames: out must be uint8 or fp32");
  if (T <= 0 || C <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0)
    return fail(VR2_ERR_ARG, "vr2_resize_frames: sizes must be positive (T=%d C=%d %dx%d -> %dx%d)", T, C, H, W, out_h, out_w);
  if (stride_t < 0 || stride_c < 0 || stride_h < 0 || stride_w < 0)
    return fail(VR2_ERR_ARG, "vr2_resize_frames: negative strides are not supported");
  cudaStream_t s = as_stream(stream_v);
  const TableLayout l = table_layout(H, W, out_h, out_w);
  char* ws = static_cast<char*>(workspace_dev);
  int* fx = reinterpret_cast<int*>(ws + l.off_fx);
  int* cx = reinterpret_cast<int*>(ws + l.off_cx);
  float* wx = reinterpret_cast<float*>(ws + l.off_wx);
  int* fy = reinterpret_cast<int*>(ws + l.off_fy);
  int* cy = reinterpret_cast<int*>(ws + l.off_cy);
  float* wy = reinterpret_cast<float*>(ws + l.off_wy);
  axis_weights_kernel<<<(out_w + 127) / 128, 128, 0, s>>>(W, out_w, l.taps_x, fx, cx, wx);
  axis_weights_kernel<<<(out_h + 127) / 128, 128, 0, s>>>(H, out_h, l.taps_y, fy, cy, wy);
  VR2_CUDA(cudaGetLastError());

  // output tile: 64 wide, 32 tall, halved until the staged window fits in shared memory
  int tile_x = 64, tile_y = 32;
  int dev = 0, smem_max = 0;
  VR2_CUDA(cudaGetDevice(&dev));
  VR2_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t budget = std::min<size_t>(static_cast<size_t>(smem_max), 100 * 1024);  // two CTAs per SM
  int cap_rows = 0, cap_cols = 0;
  size_t smem = 0;
  for (;;) {
    cap_rows = window_cap(H, out_h, tile_y);
    cap_cols = static_cast<int>(align_up(window_cap(W, out_w, tile_x), 4));
    smem = tile_smem(l, tile_x, tile_y, cap_rows, cap_cols);
    if (smem <= budget) break;
    if (tile_y > 1 && (tile_y >= tile_x / 2 || tile_x == 1)) tile_y /= 2;
    else if (tile_x > 1) tile_x /= 2;
    else if (smem <= static_cast<size_t>(smem_max)) break;
    else return fail(VR2_ERR_ARG, "vr2_resize_frames: %dx%d -> %dx%d needs %zu bytes of shared memory for a single output pixel", H, W,
                     out_h, out_w, smem);
  }
  ResizeArgs a;
  a.src = static_cast<const uint8_t*>(frames_dev);
  a.src_end = a.src + (T - 1) * stride_t + (C - 1) * stride_c + (H - 1) * stride_h + (W - 1) * stride_w + 1;
  a.st = stride_t; a.sc = stride_c; a.sh = stride_h; a.sw = stride_w;
  a.C = C; a.H = H; a.W = W; a.oh = out_h; a.ow = out_w;
  a.tx = AxisTable{fx, cx, wx, l.taps_x};
  a.ty = AxisTable{fy, cy, wy, l.taps_y};
  a.tile_x = tile_x; a.tile_y = tile_y; a.cap_cols = cap_cols; a.cap_rows = cap_rows;
  const int planes = T * C;
  if (planes > 65535) return fail(VR2_ERR_ARG, "vr2_resize_frames: T*C = %d planes exceeds 65535 per call", planes);
  dim3 grid((out_w + tile_x - 1) / tile_x, (out_h + tile_y - 1) / tile_y, planes);
  if (out_dtype == VR2_U8) {
    VR2_CUDA(cudaFuncSetAttribute(resize_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resize_kernel<uint8_t><<<grid, kThreads, smem, s>>>(a, static_cast<uint8_t*>(out_dev));
  } else {
    VR2_CUDA(cudaFuncSetAttribute(resize_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resize_kernel<float><<<grid, kThreads, smem, s>>>(a, static_cast<float*>(out_dev));
  }
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}
