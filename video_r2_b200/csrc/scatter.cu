// Placing the visual tokens for the decoder: the caller side of the tower,
//   mask = input_ids == video_token_id;  inputs_embeds = inputs_embeds.masked_scatter(mask[..., None], video_embeds)
// (Video-R2 train/src/train/monkey_patch_forward.py:385-406; HF modeling_qwen2_5_vl.py does the same in its stock
// forward).  masked_scatter consumes the rows of video_embeds in the flattened order of the True positions, so the
// work is (1) the ordered list of positions of one token id — a stream compaction — and (2) a row scatter, which the
// tower fuses into the merger's last GEMM store (vr2_tower_forward_scatter) or runs stand-alone with a dtype cast for
// cached embeddings (vr2_scatter_rows).  Integer / byte work: bit-exact, HBM-bound, coalesced 16-byte accesses.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "ptx.cuh"

namespace vr2 {

namespace {

constexpr int kScanBlock = 1024;   // elements per block of the compaction (one per thread)

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
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
    int s = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  const int base = warp > 0 ? warp_sums[warp - 1] : 0;
  *total = warp_sums[31];
  __syncthreads();
  return base + x - v;
}

// pass 1: matches per block
__global__ void __launch_bounds__(kScanBlock)
count_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t token, int* __restrict__ block_counts) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
  const int hit = (i < n && ids[i] == token) ? 1 : 0;
  const int c = __syncthreads_count(hit);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// pass 2 (one block): exclusive scan of the block counts in place, total -> counts[n_blocks]
__global__ void __launch_bounds__(kScanBlock)
scan_counts_kernel(int* __restrict__ counts, int n_blocks) {
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const int v = i < n_blocks ? counts[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    if (i < n_blocks) counts[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[n_blocks] = carry;
}

// pass 3: positions[k] = flat index of the k-th match
__global__ void __launch_bounds__(kScanBlock)
positions_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t token, const int* __restrict__ block_offsets,
                 int32_t* __restrict__ positions, int64_t capacity) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kScanBlock + threadIdx.x;
  const int hit = (i < n && ids[i] == token) ? 1 : 0;
  int total;
  const int ex = block_exclusive_scan(hit, &total);
  const int64_t k = static_cast<int64_t>(block_offsets[blockIdx.x]) + ex;
  if (hit && k < capacity) positions[k] = static_cast<int32_t>(i);
}

// composed[p] = dest_rows[out_map[p]]: window-order token p of the merger lands where the original-order token
// out_map[p] belongs in the destination
__global__ void compose_map_kernel(const int32_t* __restrict__ out_map, const int32_t* __restrict__ dest_rows,
                                   int32_t* __restrict__ composed, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) composed[i] = dest_rows[out_map[i]];
}

template <typename Src, typename Dst>
__global__ void scatter_rows_kernel(const Src* __restrict__ src, Dst* __restrict__ dst,
                                    const int32_t* __restrict__ dest_rows, int64_t n_rows, int64_t row_elems,
                                    int64_t ld_dst) {
  // 8 elements per thread: 16-byte loads for bf16, 2 x 16 for fp32
  const int64_t vec_per_row = row_elems / 8;
  const int64_t total = n_rows * vec_per_row;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = e / vec_per_row;
    const int64_t v = e - row * vec_per_row;
    const int32_t dr = dest_rows[row];
    if (dr < 0) continue;   // no destination (fewer matches than rows): skipped, the caller reports the mismatch
    const Src* s = src + row * row_elems + v * 8;
    Dst* d = dst + static_cast<int64_t>(dr) * ld_dst + v * 8;
    float f[8];
    if constexpr (sizeof(Src) == 2) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(s));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
      }
      if constexpr (sizeof(Dst) == 2) {   // same dtype: a pure byte copy
        *reinterpret_cast<uint4*>(d) = u;
        continue;
      }
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(s));
      const float4 b = __ldg(reinterpret_cast<const float4*>(s) + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    if constexpr (sizeof(Dst) == 2) {
      *reinterpret_cast<uint4*>(d) =
          make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(f[0], f[1], f[2], f[3]);
      *(reinterpret_cast<float4*>(d) + 1) = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
}

}  // namespace

cudaError_t compose_map(const int32_t* out_map, const int32_t* dest_rows, int32_t* composed, int64_t n,
                        cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  compose_map_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(out_map, dest_rows, composed, n);
  return cudaGetLastError();
}

}  // namespace vr2

using namespace vr2;

extern "C" size_t vr2_token_positions_scratch_bytes(int64_t n) {
  if (n <= 0) return 0;
  const int64_t blocks = (n + kScanBlock - 1) / kScanBlock;
  return static_cast<size_t>(blocks + 1) * 4;
}

extern "C" int vr2_token_positions(const int64_t* input_ids_dev, int64_t n, int64_t token_id, int32_t* positions_dev,
                                   int64_t capacity, void* scratch_dev, int64_t* count_host, void* stream_v) {
  if (input_ids_dev == nullptr || n < 0 || (capacity > 0 && positions_dev == nullptr))
    return fail(VR2_ERR_ARG, "vr2_token_positions: null pointer");
  if (n > INT32_MAX) return fail(VR2_ERR_ARG, "vr2_token_positions: %lld elements exceed int32 row indexing", (long long)n);
  if (count_host != nullptr) *count_host = 0;
  cudaStream_t s = as_stream(stream_v);
  // entries past the number of matches stay -1: a consumer (vr2_tower_forward_scatter, vr2_scatter_rows) skips them
  if (capacity > 0) VR2_CUDA(cudaMemsetAsync(positions_dev, 0xFF, static_cast<size_t>(capacity) * 4, s));
  if (n == 0) return VR2_OK;
  if (scratch_dev == nullptr) return fail(VR2_ERR_ARG, "vr2_token_positions: null scratch");
  const int blocks = static_cast<int>((n + kScanBlock - 1) / kScanBlock);
  int* counts = static_cast<int*>(scratch_dev);
  count_kernel<<<blocks, kScanBlock, 0, s>>>(input_ids_dev, n, token_id, counts);
  scan_counts_kernel<<<1, kScanBlock, 0, s>>>(counts, blocks);
  positions_kernel<<<blocks, kScanBlock, 0, s>>>(input_ids_dev, n, token_id, counts, positions_dev, capacity);
  VR2_CUDA(cudaGetLastError());
  if (count_host != nullptr) return vr2_token_positions_count(scratch_dev, n, count_host, stream_v);
  return VR2_OK;
}

extern "C" int vr2_token_positions_count(const void* scratch_dev, int64_t n, int64_t* count_host, void* stream_v) {
  if (count_host == nullptr) return fail(VR2_ERR_ARG, "vr2_token_positions_count: null pointer");
  *count_host = 0;
  if (n <= 0) return VR2_OK;
  if (scratch_dev == nullptr) return fail(VR2_ERR_ARG, "vr2_token_positions_count: null scratch");
  cudaStream_t s = as_stream(stream_v);
  const int64_t blocks = (n + kScanBlock - 1) / kScanBlock;
  int total = 0;
  VR2_CUDA(cudaMemcpyAsync(&total, static_cast<const int*>(scratch_dev) + blocks, 4, cudaMemcpyDeviceToHost, s));
  VR2_CUDA(cudaStreamSynchronize(s));   // the caller compares the count with the number of features (ValueError in HF)
  *count_host = total;
  return VR2_OK;
}

extern "C" int vr2_scatter_rows(const void* src_dev, int32_t src_dtype, void* dst_dev, int32_t dst_dtype,
                                const int32_t* dest_rows_dev, int64_t n_rows, int64_t row_elems, int64_t ld_dst,
                                void* stream_v) {
  if (n_rows == 0) return VR2_OK;
  if (src_dev == nullptr || dst_dev == nullptr || dest_rows_dev == nullptr || n_rows < 0)
    return fail(VR2_ERR_ARG, "vr2_scatter_rows: null pointer");
  if ((src_dtype != VR2_F32 && src_dtype != VR2_BF16) || (dst_dtype != VR2_F32 && dst_dtype != VR2_BF16))
    return fail(VR2_ERR_ARG, "vr2_scatter_rows: dtypes must be fp32 or bf16");
  if (row_elems % 8 != 0 || ld_dst % 8 != 0) return fail(VR2_ERR_ARG, "vr2_scatter_rows: row length / pitch must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(src_dev) & 15) || (reinterpret_cast<uintptr_t>(dst_dev) & 15))
    return fail(VR2_ERR_ARG, "vr2_scatter_rows: buffers must be 16-byte aligned");
  cudaStream_t s = as_stream(stream_v);
  const int64_t total = n_rows * (row_elems / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const unsigned g = static_cast<unsigned>(blocks);
  const bool sb = src_dtype == VR2_BF16, db = dst_dtype == VR2_BF16;
  if (sb && db)
    scatter_rows_kernel<<<g, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src_dev), static_cast<__nv_bfloat16*>(dst_dev), dest_rows_dev, n_rows, row_elems, ld_dst);
  else if (sb)
    scatter_rows_kernel<<<g, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src_dev), static_cast<float*>(dst_dev), dest_rows_dev, n_rows, row_elems, ld_dst);
  else if (db)
    scatter_rows_kernel<<<g, 256, 0, s>>>(static_cast<const float*>(src_dev), static_cast<__nv_bfloat16*>(dst_dev), dest_rows_dev, n_rows, row_elems, ld_dst);
  else
    scatter_rows_kernel<<<g, 256, 0, s>>>(static_cast<const float*>(src_dev), static_cast<float*>(dst_dev), dest_rows_dev, n_rows, row_elems, ld_dst);
  VR2_CUDA(cudaGetLastError());
  return VR2_OK;
}
