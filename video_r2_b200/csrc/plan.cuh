// Index plan: every table the tower derives from grid_thw alone (internal view).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

struct vr2_plan {
  int device = 0;
  int merge = 2;         // spatial_merge_size
  int win_tokens = 4;    // window edge in merged tokens (window_size / merge / patch)
  int half_dim = 40;     // rotary angles per patch (head_dim / 2)
  int64_t n_patches = 0;
  int64_t n_tokens = 0;
  int64_t n_windows = 0; // non-empty windows
  int64_t n_slices = 0;
  std::vector<int64_t> grid;  // host copy [n_items*3]

  // device tables (one allocation, carved)
  void* blob = nullptr;
  int64_t* window_index = nullptr;   // [tokens]
  int64_t* reverse_index = nullptr;  // [tokens]
  int32_t* row_map = nullptr;        // [patches] original patch row -> window-order row
  int32_t* out_map = nullptr;        // [tokens]  window-order token -> original token (== window_index)
  int32_t* pos_hw = nullptr;         // [patches,2] original order
  float* cos_tab = nullptr;          // [patches, half_dim] window order
  float* sin_tab = nullptr;
  int32_t* pos_win = nullptr;        // [patches] window order: hpos | wpos << 16
  float* pos_tab = nullptr;          // [n_pos, half_dim]: cos(p * f_k) for k < half_dim/2, then sin(p * f_k)
  int n_pos = 0;                     // max(h, w) over the clips (patch units)
  int32_t* cu_window = nullptr;      // [windows+1]
  int32_t* cu_seqlens = nullptr;     // [slices+1]
  int2* seg_window = nullptr;        // [patches] (first, past-last) row of the window a patch row belongs to
  int2* seg_full = nullptr;          // [patches] ... of its temporal slice (both attention kernels)
};
