// The vision tower behind the C ABI: weight ingestion (HF state_dict names ->
// library-owned, repacked bf16 / fp32 storage) and the forward schedule
// (HF modeling_qwen2_5_vl.py:455-518) expressed as 5 launches per block:
//
//   qkv   GEMM  A = bf16(x * w_norm1)   epilogue: * rstd(x), + bias, 2-D rotary      -> qkv  bf16 [N,3840]
//   attn        varlen flash attention over window / slice segments                  -> att  bf16 [N,1280]
//   proj  GEMM  epilogue: x += acc + bias; xb = bf16(x * w_norm2); partial sum x^2   -> x fp32, xb, ssq
//   gate|up GEMM (interleaved 128-column groups) epilogue: * rstd, + bias, SiLU*up   -> h    bf16 [N,3456]
//   down  GEMM  epilogue: x += acc + bias; xb = bf16(x * w_norm1[next]); ssq
//
// RMSNorm never runs as a kernel inside the blocks: x * rsqrt(mean x^2 + eps) * w
// feeds a linear layer, so the row scale rstd commutes with the GEMM and is
// applied to the fp32 accumulator, while w is multiplied in (fp32) before the
// single bf16 rounding of the operand.  The residual stream x stays fp32 in HBM.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "attention.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "plan.cuh"

namespace vr2 {

std::string& last_error_ref() {
  thread_local std::string s;
  return s;
}

namespace {

struct Tensor {
  void* ptr = nullptr;
  bool set = false;
};

struct Layer {
  Tensor norm1_w, norm2_w;        // fp32 [hidden]
  Tensor qkv_w, qkv_b;            // bf16 [3h, h], fp32 [3h]
  Tensor proj_w, proj_b;          // bf16 [h, h], fp32 [h]
  Tensor gu_w, gu_b;              // bf16 [2*ipad, h] interleaved per 128, fp32 [2*ipad]
  bool gate_w_set = false, up_w_set = false, gate_b_set = false, up_b_set = false;
  Tensor down_w, down_b;          // bf16 [h, ipad], fp32 [h]
  Tensor qkv_h, proj_h, gu_h, down_h;   // fp16 copies of the four matrices (VR2_PRECISION_F16), made on first use
};

}  // namespace
}  // namespace vr2

struct vr2_tower {
  vr2_tower_config cfg;
  int device = 0;
  int num_sms = 148;
  int head_dim = 80;
  int ipad = 3456;        // intermediate padded to a multiple of 128
  int patch_k = 1176;     // in_channels * temporal_patch * patch * patch
  int merge_dim = 5120;   // hidden * merge^2
  bool finalized = false;
  int precision = VR2_PRECISION_BF16;
  std::vector<vr2::Layer> layers;
  vr2::Tensor patch_w;             // bf16 [h, patch_k]
  vr2::Tensor lnq_w;               // fp32 [h]
  vr2::Tensor fc1_w, fc1_b;        // bf16 [md, md], fp32 [md]
  vr2::Tensor fc2_w, fc2_b;        // bf16 [out, md], fp32 [out]
  vr2::Tensor fc1_h, fc2_h;        // fp16 copies (VR2_PRECISION_F16)
  bool f16_ready = false;          // the fp16 copies reflect the current weights
  // f16 mode's range guard: device counters of fp16 conversions that saturated, one per site
  //   [0] xb (the residual-stream operand x * w_norm)   [1] q/k/v   [2] h (SwiGLU product)   [3] merger (ln_q output, fc1 hidden)
  unsigned int* sat = nullptr;
  std::vector<float> norm_wmax2;   // [2 * depth + 1] largest squared weight of norm1[l], norm2[l], ..., ln_q (finalize)
  // f16 mode's range shifts: xb is stored times 2^-shift_xb (its consumers multiply rstd by 2^shift_xb), h times
  // 2^-shift_h (the down GEMM's epilogue multiplies its accumulator by 2^shift_h).  Powers of two: exact.  They are
  // applied through pre-scaled COPIES of the norm weights / up-biases, so the kernels spend no instruction on them.
  int shift_xb = 0, shift_h = 0;
  std::vector<std::vector<float>> norm_host;   // [2 * depth + 1][hidden] host copies (finalize)
  std::vector<std::vector<float>> gub_host;    // [depth][2 * ipad]
  std::vector<vr2::Tensor> norm_s, gub_s;      // device copies scaled for the current shifts (set_f16_shift)
  std::vector<void*> allocs;
  // optional per-kind CUDA-event timing (vr2_tower_profile)
  uint32_t prof_mask = 0;
  std::vector<cudaEvent_t> prof_events;  // pairs (begin, end)
  std::vector<int> prof_kind;            // kind of pair i
  size_t prof_used = 0;                  // pairs recorded since the last reset
  int debug_stop = 0;                    // tests: return from the forward after this many stages (vr2_tower_set_debug_stop)
};

using namespace vr2;

namespace {

int alloc_zero(vr2_tower* t, Tensor* dst, size_t bytes) {
  void* p = nullptr;
  VR2_CUDA(cudaMalloc(&p, bytes));
  VR2_CUDA(cudaMemset(p, 0, bytes));
  t->allocs.push_back(p);
  dst->ptr = p;
  return VR2_OK;
}

bool is_full_layer(const vr2_tower* t, int l) {
  for (int i = 0; i < t->cfg.n_fullatt; ++i)
    if (t->cfg.fullatt_block_indexes[i] == l) return true;
  return false;
}

// fp16 copies of the GEMM weights, converted on the forward's stream the first time the f16 mode runs (and again after
// a weight changes).  tcgen05 kind::f16 rejects mixed fp16 x bf16 operands (illegal instruction), so the f16
// activation mode needs fp16 weights; bf16 -> fp16 is exact inside fp16's normal range.
int ensure_f16_weights(vr2_tower* t, cudaStream_t s) {
  if (t->f16_ready) return VR2_OK;
  const size_t h = t->cfg.hidden_size, ip = t->ipad, md = t->merge_dim, oh = t->cfg.out_hidden_size;
  auto conv = [&](const Tensor& src, Tensor* dst, size_t n) -> int {
    if (dst->ptr == nullptr) {
      void* p = nullptr;
      VR2_CUDA(cudaMalloc(&p, n * 2));
      t->allocs.push_back(p);
      dst->ptr = p;
    }
    VR2_CUDA(cast_bf16_to_f16(static_cast<const __nv_bfloat16*>(src.ptr), dst->ptr, static_cast<int64_t>(n), s));
    return VR2_OK;
  };
  int rc;
  if ((rc = conv(t->fc1_w, &t->fc1_h, md * md)) != VR2_OK) return rc;
  if ((rc = conv(t->fc2_w, &t->fc2_h, oh * md)) != VR2_OK) return rc;
  for (auto& L : t->layers) {
    if ((rc = conv(L.qkv_w, &L.qkv_h, 3 * h * h)) != VR2_OK) return rc;
    if ((rc = conv(L.proj_w, &L.proj_h, h * h)) != VR2_OK) return rc;
    if ((rc = conv(L.gu_w, &L.gu_h, 2 * ip * h)) != VR2_OK) return rc;
    if ((rc = conv(L.down_w, &L.down_h, h * ip)) != VR2_OK) return rc;
  }
  t->f16_ready = true;
  return VR2_OK;
}

constexpr size_t kMaxProfPairs = 8192;

// Records an event pair around the launches issued while it is alive, when `kind` is selected.
struct ProfScope {
  vr2_tower* t;
  cudaStream_t s;
  cudaEvent_t end = nullptr;
  ProfScope(vr2_tower* tower, int kind, cudaStream_t stream) : t(tower), s(stream) {
    if (!(t->prof_mask & (1u << kind)) || t->prof_used >= kMaxProfPairs) return;
    const size_t i = t->prof_used;
    if (t->prof_events.size() < 2 * (i + 1)) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      t->prof_events.push_back(a);
      t->prof_events.push_back(b);
      t->prof_kind.push_back(kind);
    }
    t->prof_kind[i] = kind;
    cudaEventRecord(t->prof_events[2 * i], s);
    end = t->prof_events[2 * i + 1];
    t->prof_used = i + 1;
  }
  ~ProfScope() {
    if (end != nullptr) cudaEventRecord(end, s);
  }
};

}  // namespace

extern "C" const char* vr2_last_error(void) { return last_error_ref().c_str(); }
extern "C" int vr2_abi_version(void) { return 1; }

extern "C" int vr2_tower_create(const vr2_tower_config* cfg, int32_t device, vr2_tower** out) {
  if (cfg == nullptr || out == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_create: null argument");
  if (cfg->depth <= 0 || cfg->num_heads <= 0 || cfg->hidden_size % cfg->num_heads != 0)
    return fail(VR2_ERR_ARG, "vr2_tower_create: bad depth/heads/hidden");
  const int head_dim = cfg->hidden_size / cfg->num_heads;
  if (head_dim != 80) return fail(VR2_ERR_ARG, "vr2_tower_create: head_dim %d unsupported (kernels are built for 80)", head_dim);
  if (cfg->hidden_size % 160 != 0 || (3 * cfg->hidden_size) % 240 != 0 || (cfg->hidden_size / 40) % 4 != 0)
    return fail(VR2_ERR_ARG, "vr2_tower_create: hidden_size %d must be a multiple of 1280", cfg->hidden_size);
  if (cfg->spatial_merge_size != 2) return fail(VR2_ERR_ARG, "vr2_tower_create: spatial_merge_size must be 2");
  if (cfg->out_hidden_size % 32 != 0) return fail(VR2_ERR_ARG, "vr2_tower_create: out_hidden_size must be a multiple of 32");
  if (cfg->n_fullatt < 0 || cfg->n_fullatt > 8) return fail(VR2_ERR_ARG, "vr2_tower_create: n_fullatt out of range");
  const int patch_k = cfg->in_channels * cfg->temporal_patch_size * cfg->patch_size * cfg->patch_size;
  if (patch_k % 8 != 0) return fail(VR2_ERR_ARG, "vr2_tower_create: patch row of %d elements is not 16-byte aligned", patch_k);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= device)
    return fail(VR2_ERR_CUDA, "vr2_tower_create: CUDA device %d not available (%s); there is no CPU path", device,
                e != cudaSuccess ? cudaGetErrorString(e) : "device count too small");
  // allocations below go to `device`; the caller's current device is restored on every exit path
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } guard;
  VR2_CUDA(cudaGetDevice(&guard.prev));
  VR2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VR2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(VR2_ERR_CUDA, "vr2_tower_create: device is sm_%d%d; this library contains sm_100a code only", prop.major, prop.minor);

  auto* t = new vr2_tower();
  t->cfg = *cfg;
  t->device = device;
  t->num_sms = prop.multiProcessorCount;
  t->head_dim = head_dim;
  t->ipad = (cfg->intermediate_size + 127) / 128 * 128;
  t->patch_k = patch_k;
  t->merge_dim = cfg->hidden_size * cfg->spatial_merge_size * cfg->spatial_merge_size;
  t->layers.resize(cfg->depth);
  const size_t h = cfg->hidden_size, ip = t->ipad, md = t->merge_dim, oh = cfg->out_hidden_size;
  int rc = VR2_OK;
  auto A = [&](Tensor* x, size_t bytes) { if (rc == VR2_OK) rc = alloc_zero(t, x, bytes); };
  A(&t->patch_w, h * patch_k * 2);
  A(&t->lnq_w, h * 4);
  A(&t->fc1_w, md * md * 2);
  A(&t->fc1_b, md * 4);
  A(&t->fc2_w, oh * md * 2);
  A(&t->fc2_b, oh * 4);
  for (auto& L : t->layers) {
    A(&L.norm1_w, h * 4);
    A(&L.norm2_w, h * 4);
    A(&L.qkv_w, 3 * h * h * 2);
    A(&L.qkv_b, 3 * h * 4);
    A(&L.proj_w, h * h * 2);
    A(&L.proj_b, h * 4);
    A(&L.gu_w, 2 * ip * h * 2);
    A(&L.gu_b, 2 * ip * 4);
    A(&L.down_w, h * ip * 2);
    A(&L.down_b, h * 4);
  }
  if (rc == VR2_OK) {
    Tensor c;
    rc = alloc_zero(t, &c, 64);   // VR2_SAT_SITES counters
    t->sat = static_cast<unsigned int*>(c.ptr);
  }
  if (rc != VR2_OK) {
    vr2_tower_destroy(t);
    return rc;
  }
  *out = t;
  return VR2_OK;
}

extern "C" void vr2_tower_destroy(vr2_tower* t) {
  if (t == nullptr) return;
  for (void* p : t->allocs) cudaFree(p);
  for (cudaEvent_t e : t->prof_events) cudaEventDestroy(e);
  delete t;
}

extern "C" int vr2_tower_set_weight(vr2_tower* t, const char* name, const void* data, int32_t dtype,
                                    const int64_t* shape, int32_t ndim, void* stream_v) {
  if (t == nullptr || name == nullptr || data == nullptr || shape == nullptr)
    return fail(VR2_ERR_ARG, "vr2_tower_set_weight: null argument");
  t->f16_ready = false;
  if (dtype != VR2_F32 && dtype != VR2_BF16)
    return fail(VR2_ERR_ARG, "vr2_tower_set_weight(%s): dtype %d unsupported (fp32/bf16)", name, dtype);
  cudaStream_t s = as_stream(stream_v);
  const bool sb = dtype == VR2_BF16;
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  const int64_t h = t->cfg.hidden_size, inter = t->cfg.intermediate_size, ip = t->ipad, md = t->merge_dim,
                oh = t->cfg.out_hidden_size;
  const int64_t huge = int64_t(1) << 40;

  // dst rows = src rows unless (blk, stride, off) say otherwise
  auto mat = [&](Tensor* dst, bool dst_bf16, int64_t rows, int64_t cols, int64_t ld, int64_t blk, int64_t stride,
                 int64_t off) -> int {
    if (numel != rows * cols)
      return fail(VR2_ERR_ARG, "vr2_tower_set_weight(%s): expected %lld x %lld = %lld elements, got %lld", name,
                  (long long)rows, (long long)cols, (long long)(rows * cols), (long long)numel);
    VR2_CUDA(repack_matrix(data, sb, rows, cols, dst->ptr, dst_bf16, ld, blk, stride, off, s));
    dst->set = true;
    return VR2_OK;
  };
  auto W = [&](Tensor* dst, int64_t rows, int64_t cols, int64_t ld) { return mat(dst, true, rows, cols, ld, huge, 0, 0); };
  auto V = [&](Tensor* dst, int64_t n) { return mat(dst, false, 1, n, n, huge, 0, 0); };

  const std::string n(name);
  if (n == "patch_embed.proj.weight") return W(&t->patch_w, h, t->patch_k, t->patch_k);
  if (n == "merger.ln_q.weight") return V(&t->lnq_w, h);
  if (n == "merger.mlp.0.weight") return W(&t->fc1_w, md, md, md);
  if (n == "merger.mlp.0.bias") return V(&t->fc1_b, md);
  if (n == "merger.mlp.2.weight") return W(&t->fc2_w, oh, md, md);
  if (n == "merger.mlp.2.bias") return V(&t->fc2_b, oh);
  if (n.rfind("blocks.", 0) == 0) {
    const size_t dot = n.find('.', 7);
    if (dot == std::string::npos) return fail(VR2_ERR_ARG, "vr2_tower_set_weight: cannot parse '%s'", name);
    const int l = atoi(n.substr(7, dot - 7).c_str());
    if (l < 0 || l >= t->cfg.depth) return fail(VR2_ERR_ARG, "vr2_tower_set_weight: layer %d out of range in '%s'", l, name);
    Layer& L = t->layers[l];
    const std::string k = n.substr(dot + 1);
    if (k == "norm1.weight") return V(&L.norm1_w, h);
    if (k == "norm2.weight") return V(&L.norm2_w, h);
    if (k == "attn.qkv.weight") return W(&L.qkv_w, 3 * h, h, h);
    if (k == "attn.qkv.bias") return V(&L.qkv_b, 3 * h);
    if (k == "attn.proj.weight") return W(&L.proj_w, h, h, h);
    if (k == "attn.proj.bias") return V(&L.proj_b, h);
    // gate rows g -> tile (g/128), columns [0,128); up rows u -> same tile, columns [128,256)
    if (k == "mlp.gate_proj.weight") { int rc = mat(&L.gu_w, true, inter, h, h, 128, 256, 0); L.gate_w_set = rc == VR2_OK; return rc; }
    if (k == "mlp.up_proj.weight") { int rc = mat(&L.gu_w, true, inter, h, h, 128, 256, 128); L.up_w_set = rc == VR2_OK; return rc; }
    if (k == "mlp.gate_proj.bias") { int rc = mat(&L.gu_b, false, inter, 1, 1, 128, 256, 0); L.gate_b_set = rc == VR2_OK; return rc; }
    if (k == "mlp.up_proj.bias") { int rc = mat(&L.gu_b, false, inter, 1, 1, 128, 256, 128); L.up_b_set = rc == VR2_OK; return rc; }
    if (k == "mlp.down_proj.weight") return W(&L.down_w, h, inter, ip);  // K padded with zeros to ipad
    if (k == "mlp.down_proj.bias") return V(&L.down_b, h);
  }
  return fail(VR2_ERR_ARG, "vr2_tower_set_weight: unknown tensor '%s'", name);
}

extern "C" int vr2_tower_finalize(vr2_tower* t) {
  if (t == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_finalize: null tower");
  std::string missing;
  auto need = [&](bool ok, const std::string& nm) { if (!ok) missing += (missing.empty() ? "" : ", ") + nm; };
  need(t->patch_w.set, "patch_embed.proj.weight");
  need(t->lnq_w.set, "merger.ln_q.weight");
  need(t->fc1_w.set, "merger.mlp.0.weight");
  need(t->fc1_b.set, "merger.mlp.0.bias");
  need(t->fc2_w.set, "merger.mlp.2.weight");
  need(t->fc2_b.set, "merger.mlp.2.bias");
  for (int l = 0; l < t->cfg.depth; ++l) {
    Layer& L = t->layers[l];
    const std::string p = "blocks." + std::to_string(l) + ".";
    need(L.norm1_w.set, p + "norm1.weight");
    need(L.norm2_w.set, p + "norm2.weight");
    need(L.qkv_w.set, p + "attn.qkv.weight");
    need(L.qkv_b.set, p + "attn.qkv.bias");
    need(L.proj_w.set, p + "attn.proj.weight");
    need(L.proj_b.set, p + "attn.proj.bias");
    need(L.gate_w_set, p + "mlp.gate_proj.weight");
    need(L.up_w_set, p + "mlp.up_proj.weight");
    need(L.gate_b_set, p + "mlp.gate_proj.bias");
    need(L.up_b_set, p + "mlp.up_proj.bias");
    need(L.down_w.set, p + "mlp.down_proj.weight");
    need(L.down_b.set, p + "mlp.down_proj.bias");
  }
  if (!missing.empty()) {
    if (missing.size() > 600) missing = missing.substr(0, 600) + " ...";
    return fail(VR2_ERR_STATE, "vr2_tower_finalize: missing tensors: %s", missing.c_str());
  }
  // largest squared RMSNorm weight per norm vector: the bound the residual epilogues' f16 range guard uses
  {
    const size_t h = t->cfg.hidden_size;
    const int n_norm = 2 * t->cfg.depth + 1;
    t->norm_host.assign(n_norm, std::vector<float>(h));
    t->norm_wmax2.assign(n_norm, 0.f);
    auto wmax2 = [&](const Tensor& w, int i) -> int {
      VR2_CUDA(cudaMemcpy(t->norm_host[i].data(), w.ptr, h * sizeof(float), cudaMemcpyDeviceToHost));
      float m = 0.f;
      for (float v : t->norm_host[i]) m = std::max(m, v * v);
      t->norm_wmax2[i] = m;
      return VR2_OK;
    };
    int rc;
    t->gub_host.assign(t->cfg.depth, std::vector<float>(2 * static_cast<size_t>(t->ipad)));
    for (int l = 0; l < t->cfg.depth; ++l) {
      if ((rc = wmax2(t->layers[l].norm1_w, 2 * l)) != VR2_OK) return rc;
      if ((rc = wmax2(t->layers[l].norm2_w, 2 * l + 1)) != VR2_OK) return rc;
      VR2_CUDA(cudaMemcpy(t->gub_host[l].data(), t->layers[l].gu_b.ptr, t->gub_host[l].size() * sizeof(float),
                          cudaMemcpyDeviceToHost));
    }
    if ((rc = wmax2(t->lnq_w, 2 * t->cfg.depth)) != VR2_OK) return rc;
    if (t->shift_xb != 0 || t->shift_h != 0) {   // weights changed under non-zero shifts: rebuild the scaled copies
      const int sx = t->shift_xb, sh = t->shift_h;
      t->shift_xb = t->shift_h = 0;
      t->finalized = true;
      if ((rc = vr2_tower_set_f16_shift(t, sx, sh)) != VR2_OK) return rc;
    }
  }
  t->finalized = true;
  return VR2_OK;
}

namespace {

struct Workspace {
  float* x;
  __nv_bfloat16* xb;
  __nv_bfloat16* qkv;   // also the merger's normalised input [N/4, 4h]
  __nv_bfloat16* att;
  __nv_bfloat16* hbuf;  // also the merger's hidden [N/4, 4h]
  float* ssq;
  __nv_bfloat16* pix;
  int32_t* cmap;        // [N/4] merger store map composed with the caller's destination rows (forward_scatter)
  // VR2_PRECISION_SPLIT: rounding residuals of the buffers above (same layouts), else null
  __nv_bfloat16* xb_lo = nullptr;
  __nv_bfloat16* qkv_lo = nullptr;
  __nv_bfloat16* att_lo = nullptr;
  __nv_bfloat16* hbuf_lo = nullptr;
  __nv_bfloat16* pix_lo = nullptr;
  size_t bytes;
};

Workspace carve_ws(const vr2_tower* t, int64_t n, void* base) {
  auto rnd = [](size_t b) { return (b + 1023) & ~static_cast<size_t>(1023); };
  const size_t h = t->cfg.hidden_size, N = static_cast<size_t>(n);
  char* cur = static_cast<char*>(base);
  Workspace w;
  auto take = [&](size_t bytes) { char* p = cur; cur += rnd(bytes); return p; };
  w.x = reinterpret_cast<float*>(take(N * h * 4));
  w.xb = reinterpret_cast<__nv_bfloat16*>(take(N * h * 2));
  w.qkv = reinterpret_cast<__nv_bfloat16*>(take(N * 3 * h * 2));
  w.att = reinterpret_cast<__nv_bfloat16*>(take(N * h * 2));
  w.hbuf = reinterpret_cast<__nv_bfloat16*>(take(N * std::max<size_t>(t->ipad, h) * 2));
  w.ssq = reinterpret_cast<float*>(take(N * (h / 40) * 4));
  w.pix = reinterpret_cast<__nv_bfloat16*>(take(N * t->patch_k * 2));
  w.cmap = reinterpret_cast<int32_t*>(take(N));   // N/4 tokens x 4 bytes
  if (t->precision == VR2_PRECISION_SPLIT) {
    w.xb_lo = reinterpret_cast<__nv_bfloat16*>(take(N * h * 2));
    w.qkv_lo = reinterpret_cast<__nv_bfloat16*>(take(N * 3 * h * 2));
    w.att_lo = reinterpret_cast<__nv_bfloat16*>(take(N * h * 2));
    w.hbuf_lo = reinterpret_cast<__nv_bfloat16*>(take(N * std::max<size_t>(t->ipad, h) * 2));
    w.pix_lo = reinterpret_cast<__nv_bfloat16*>(take(N * t->patch_k * 2));
  }
  w.bytes = static_cast<size_t>(cur - static_cast<char*>(base));
  return w;
}

}  // namespace

extern "C" size_t vr2_tower_workspace_bytes(const vr2_tower* t, int64_t n_patches) {
  if (t == nullptr || n_patches <= 0) return 0;
  return carve_ws(t, n_patches, nullptr).bytes;
}

extern "C" int64_t vr2_tower_launch_count(const vr2_tower* t, const vr2_plan* plan, int32_t pixels_dtype) {
  if (t == nullptr || plan == nullptr) return -1;
  return (pixels_dtype == VR2_F32 ? 1 : 0) + 1 + 5 * static_cast<int64_t>(t->cfg.depth) + 3;  // memset is not a kernel
}

namespace vr2 {
cudaError_t compose_map(const int32_t* out_map, const int32_t* dest_rows, int32_t* composed, int64_t n,
                        cudaStream_t stream);
}

// dest_rows == nullptr: merged token k -> out[k, :] (row pitch out_hidden).  Otherwise token k -> out[dest_rows[k], :]
// with row pitch ld_out: the masked_scatter of the caller fused into the merger's last store.
// multi != nullptr: the result (bf16) is stored to the n_multi buffers multi[0..] instead of `out` (vr2_tower_forward_multi).
static int tower_forward_impl(vr2_tower* t, const vr2_plan* plan, const void* pixels, int32_t pixels_dtype, void* out,
                              int32_t out_dtype, const int32_t* dest_rows, int64_t ld_out, void* workspace,
                              size_t workspace_bytes, void* stream_v, void* const* multi = nullptr, int n_multi = 0) {
  if (t == nullptr || plan == nullptr || pixels == nullptr || out == nullptr || workspace == nullptr)
    return fail(VR2_ERR_ARG, "vr2_tower_forward: null argument");
  if (!t->finalized) return fail(VR2_ERR_STATE, "vr2_tower_forward: weights not finalized");
  if (pixels_dtype != VR2_F32 && pixels_dtype != VR2_BF16) return fail(VR2_ERR_ARG, "vr2_tower_forward: pixels dtype %d unsupported", pixels_dtype);
  if (out_dtype != VR2_F32 && out_dtype != VR2_BF16) return fail(VR2_ERR_ARG, "vr2_tower_forward: out dtype %d unsupported", out_dtype);
  if (plan->half_dim * 2 != t->head_dim || plan->merge != t->cfg.spatial_merge_size)
    return fail(VR2_ERR_STATE, "vr2_tower_forward: plan was built for a different head_dim / merge size");
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail(VR2_ERR_ARG, "vr2_tower_forward: workspace must be 1024-byte aligned");
  if ((reinterpret_cast<uintptr_t>(pixels) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(VR2_ERR_ARG, "vr2_tower_forward: pixels/out must be 16-byte aligned");
  const int64_t N = plan->n_patches;
  const Workspace w = carve_ws(t, N, workspace);
  if (w.bytes > workspace_bytes)
    return fail(VR2_ERR_ARG, "vr2_tower_forward: workspace %zu < required %zu bytes", workspace_bytes, w.bytes);
  cudaStream_t s = as_stream(stream_v);
  const int h = t->cfg.hidden_size;
  // ---- precision of this call
  int rc0 = VR2_OK;
  const bool split = t->precision == VR2_PRECISION_SPLIT;
  // f16 activation mode: every 16-bit activation buffer (xb, qkv, att, h, the merger's intermediates) holds fp16 instead
  // of bf16 and the GEMMs read fp16 copies of the weights (ensure_f16_weights); the patch-embed GEMM multiplies the
  // caller's bf16 pixels with the bf16 weight as before.  Conversions that saturate are counted in t->sat.
  const bool f16 = t->precision == VR2_PRECISION_F16;
  int parts = 0;  // sum-of-squares partials per row written by the most recent EPI_RESID launch
  const float eps = t->cfg.rms_norm_eps;
  const char* msg = "";
  // Consecutive launches walk the rows in opposite directions (GEMM_REVERSE_M / the attention kernel's `reverse`):
  // every launch then starts on the rows its producer wrote last — the ones still in the 126 MB L2 — instead of on the
  // rows that were evicted longest ago.  Same tiles and arithmetic either way (VR2_ALTERNATE=0 turns it off for A/B).
  const char* alt_env = getenv("VR2_ALTERNATE");   // read per forward: tests toggle it
  const bool alternate = alt_env == nullptr || alt_env[0] != '0';
  int n_launch = 0;
  // tests (vr2_tower_set_debug_stop): the forward returns after `debug_stop` stages — patch embed, then qkv, attention,
  // proj, gate|up, down of every block — leaving the workspace as that stage wrote it
  int n_stage = 0;
  auto stage_done = [&]() { return t->debug_stop > 0 && ++n_stage >= t->debug_stop; };
  auto G = [&](GemmArgs& a, int kind, const char* what) -> int {
    ProfScope prof(t, kind, s);
    if (alternate && (n_launch++ & 1)) a.flags |= GEMM_REVERSE_M;
    const int site = a.epilogue == EPI_QKV ? 1 : a.epilogue == EPI_SWIGLU ? 2 : a.epilogue == EPI_RESID ? 0 : 3;
    a.sat = f16 ? t->sat + site : nullptr;
    if (a.epilogue == EPI_RESID) parts = gemm_resid_parts(a);
    cudaError_t e = gemm_launch(a, t->num_sms, s, &msg);
    if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? VR2_ERR_ARG : VR2_ERR_CUDA, "vr2_tower_forward: %s: %s %s", what, cudaGetErrorString(e), msg);
    return VR2_OK;
  };
  int rc;
  const int act_in = f16 ? GEMM_A_F16 : 0, act_out = f16 ? GEMM_OUT_F16 : 0;
  if (f16 && (rc0 = ensure_f16_weights(t, s)) != VR2_OK) return rc0;
  auto WT = [&](const Tensor& w_bf16, const Tensor& w_f16) { return static_cast<const __nv_bfloat16*>(f16 ? w_f16.ptr : w_bf16.ptr); };
  // range shifts of the f16 mode (vr2_tower_set_f16_shift): scaled norm-weight / up-bias copies, accumulator factors
  const int sx = f16 ? t->shift_xb : 0, sh = f16 ? t->shift_h : 0;
  const float xb_back = ldexpf(1.f, sx), h_back = ldexpf(1.f, sh), h_down = ldexpf(1.f, -sh);
  auto NW = [&](int i, const Tensor& plain) { return static_cast<const float*>(sx != 0 ? t->norm_s[i].ptr : plain.ptr); };
  auto NWMAX2 = [&](int i) { return t->norm_wmax2[i] * ldexpf(1.f, -2 * sx); };
  const __nv_bfloat16* pix = static_cast<const __nv_bfloat16*>(pixels);
  if (pixels_dtype == VR2_F32) {
    ProfScope prof(t, VR2_K_CAST, s);
    VR2_CUDA(cast_f32_to_bf16(static_cast<const float*>(pixels), w.pix, w.pix_lo, N * t->patch_k, s));
    pix = w.pix;
  } else if (split) {
    VR2_CUDA(cudaMemsetAsync(w.pix_lo, 0, static_cast<size_t>(N) * t->patch_k * 2, s));  // bf16 pixels: no residual
  }
  {
    GemmArgs a{};
    a.A = pix; a.A_lo = w.pix_lo; a.lda = t->patch_k; a.W = static_cast<const __nv_bfloat16*>(t->patch_w.ptr); a.ldw = t->patch_k;
    a.M = static_cast<int>(N); a.N = h; a.K = t->patch_k;
    a.epilogue = EPI_RESID; a.flags = act_out;
    a.x = w.x; a.ldx = h; a.out_bf16 = w.xb; a.out_lo = w.xb_lo; a.ld_out = h;
    a.row_map = plan->row_map;
    a.norm_w = NW(0, t->layers[0].norm1_w);
    a.sat_wmax2 = NWMAX2(0);
    a.ssq = w.ssq;
    if ((rc = G(a, VR2_K_PATCH_EMBED, "patch_embed")) != VR2_OK) return rc;
    if (stage_done()) return VR2_OK;
  }

  for (int l = 0; l < t->cfg.depth; ++l) {
    const Layer& L = t->layers[l];
    {
      GemmArgs a{};
      a.A = w.xb; a.A_lo = w.xb_lo; a.lda = h; a.W = WT(L.qkv_w, L.qkv_h); a.ldw = h;
      a.M = static_cast<int>(N); a.N = 3 * h; a.K = h;
      a.epilogue = EPI_QKV; a.flags = act_in | act_out;
      a.bias = static_cast<const float*>(L.qkv_b.ptr);
      a.out_bf16 = w.qkv; a.out_lo = w.qkv_lo; a.ld_out = 3 * h;
      a.ssq_in = w.ssq; a.ssq_parts = parts; a.inv_dim = 1.0f / h; a.eps = eps;
      a.in_scale = xb_back;
      a.cos_tab = plan->cos_tab; a.sin_tab = plan->sin_tab;
      a.pos_win = plan->pos_win; a.pos_tab = plan->pos_tab; a.n_pos = plan->n_pos;
      if ((rc = G(a, VR2_K_QKV, "qkv")) != VR2_OK) return rc;
      if (stage_done()) return VR2_OK;
    }
    {
      const bool full = is_full_layer(t, l);
      ProfScope prof(t, full ? VR2_K_ATTN_FULL : VR2_K_ATTN_WINDOW, s);
      if (split)   // hi + lo operands: 3 tcgen05 products per contraction (attention3.cu)
        VR2_CUDA(attention_split_launch(w.qkv, w.qkv_lo, w.att, w.att_lo, full ? plan->seg_full : plan->seg_window,
                                        static_cast<int>(N), t->cfg.num_heads, 3 * h, h, t->num_sms, s));
      else
        VR2_CUDA(attention_tc_launch(w.qkv, w.att, full ? plan->seg_full : plan->seg_window, static_cast<int>(N),
                                     t->cfg.num_heads, 3 * h, h, t->num_sms, s, f16, alternate && (n_launch & 1)));
      ++n_launch;
      if (stage_done()) return VR2_OK;
    }
    {
      GemmArgs a{};
      a.A = w.att; a.A_lo = w.att_lo; a.lda = h; a.W = WT(L.proj_w, L.proj_h); a.ldw = h;
      a.M = static_cast<int>(N); a.N = h; a.K = h;
      a.epilogue = EPI_RESID; a.flags = GEMM_ADD_RESID | act_in | act_out;
      a.bias = static_cast<const float*>(L.proj_b.ptr);
      a.x = w.x; a.ldx = h; a.out_bf16 = w.xb; a.out_lo = w.xb_lo; a.ld_out = h;
      a.norm_w = NW(2 * l + 1, L.norm2_w);
      a.sat_wmax2 = NWMAX2(2 * l + 1);
      a.ssq = w.ssq;
      if ((rc = G(a, VR2_K_PROJ, "proj")) != VR2_OK) return rc;
      if (stage_done()) return VR2_OK;
    }
    {
      GemmArgs a{};
      a.A = w.xb; a.A_lo = w.xb_lo; a.lda = h; a.W = WT(L.gu_w, L.gu_h); a.ldw = h;
      a.M = static_cast<int>(N); a.N = 2 * t->ipad; a.K = h;
      a.epilogue = EPI_SWIGLU; a.flags = act_in | act_out;
      a.bias = static_cast<const float*>(sh != 0 ? t->gub_s[l].ptr : L.gu_b.ptr);
      a.out_bf16 = w.hbuf; a.out_lo = w.hbuf_lo; a.ld_out = t->ipad;
      a.ssq_in = w.ssq; a.ssq_parts = parts; a.inv_dim = 1.0f / h; a.eps = eps;
      a.in_scale = xb_back; a.up_scale = h_down;
      if ((rc = G(a, VR2_K_GATE_UP, "gate_up")) != VR2_OK) return rc;
      if (stage_done()) return VR2_OK;
    }
    {
      GemmArgs a{};
      a.A = w.hbuf; a.A_lo = w.hbuf_lo; a.lda = t->ipad; a.W = WT(L.down_w, L.down_h); a.ldw = t->ipad;
      a.M = static_cast<int>(N); a.N = h; a.K = t->ipad;
      a.epilogue = EPI_RESID; a.flags = GEMM_ADD_RESID | act_in | act_out;
      a.bias = static_cast<const float*>(L.down_b.ptr);
      a.x = w.x; a.ldx = h; a.out_bf16 = w.xb; a.out_lo = w.xb_lo; a.ld_out = h;
      a.norm_w = NW(2 * l + 2, l + 1 < t->cfg.depth ? t->layers[l + 1].norm1_w : t->lnq_w);
      a.sat_wmax2 = NWMAX2(2 * l + 2);
      a.in_scale = h_back;
      a.ssq = w.ssq;
      if ((rc = G(a, VR2_K_DOWN, "down")) != VR2_OK) return rc;
      if (stage_done()) return VR2_OK;
    }
  }

  // ---- PatchMerger: ln_q per patch row, then 4 consecutive window-order rows form one merged token
  const int64_t T = plan->n_tokens;
  const int md = t->merge_dim;
  {
    ProfScope prof(t, VR2_K_MERGER_NORM, s);
    VR2_CUDA(rmsnorm_to_bf16(w.x, static_cast<const float*>(t->lnq_w.ptr), w.qkv, w.qkv_lo, N, h, eps, s, f16,
                             f16 ? t->sat + 3 : nullptr));
  }
  {
    GemmArgs a{};
    a.A = w.qkv; a.A_lo = w.qkv_lo; a.lda = md; a.W = WT(t->fc1_w, t->fc1_h); a.ldw = md;
    a.M = static_cast<int>(T); a.N = md; a.K = md;
    a.epilogue = EPI_STORE; a.flags = GEMM_GELU | act_in | act_out;
    a.bias = static_cast<const float*>(t->fc1_b.ptr);
    a.out_bf16 = w.hbuf; a.out_lo = w.hbuf_lo; a.ld_out = md;
    if ((rc = G(a, VR2_K_FC1, "merger.fc1")) != VR2_OK) return rc;
  }
  {
    GemmArgs a{};
    a.A = w.hbuf; a.A_lo = w.hbuf_lo; a.lda = md; a.W = WT(t->fc2_w, t->fc2_h); a.ldw = md;
    a.M = static_cast<int>(T); a.N = t->cfg.out_hidden_size; a.K = md;
    a.epilogue = EPI_STORE; a.flags = (out_dtype == VR2_F32 ? GEMM_OUT_F32 : 0) | act_in;   // the result itself: bf16 / fp32
    a.bias = static_cast<const float*>(t->fc2_b.ptr);
    a.out_bf16 = static_cast<__nv_bfloat16*>(out); a.out_f32 = static_cast<float*>(out);
    a.ld_out = t->cfg.out_hidden_size;
    a.row_map = plan->out_map;  // inverse window permute fused into the store (HF :512-513)
    if (multi != nullptr) {
      // sharded multi-GPU mode: every row goes to all destinations (this rank's and the peers' copies of the result)
      a.epilogue = EPI_STORE_MULTI;
      a.flags = act_in;
      a.n_multi = n_multi;
      for (int d = 0; d < n_multi; ++d) a.multi_out[d] = multi[d];
    }
    if (dest_rows != nullptr) {
      VR2_CUDA(compose_map(plan->out_map, dest_rows, w.cmap, T, s));
      a.row_map = w.cmap;
      a.ld_out = ld_out;
    }
    if ((rc = G(a, VR2_K_FC2, "merger.fc2")) != VR2_OK) return rc;
  }
  return VR2_OK;
}

extern "C" int vr2_tower_forward(vr2_tower* t, const vr2_plan* plan, const void* pixels, int32_t pixels_dtype,
                                 void* out, int32_t out_dtype, void* workspace, size_t workspace_bytes,
                                 void* stream_v) {
  return tower_forward_impl(t, plan, pixels, pixels_dtype, out, out_dtype, nullptr, 0, workspace, workspace_bytes, stream_v);
}

extern "C" int vr2_tower_forward_multi(vr2_tower* t, const vr2_plan* plan, const void* pixels, int32_t pixels_dtype,
                                       void* const* outs, int32_t n_out, void* workspace, size_t workspace_bytes,
                                       void* stream_v) {
  if (outs == nullptr || n_out < 1 || n_out > kMaxMultiOut)
    return fail(VR2_ERR_ARG, "vr2_tower_forward_multi: 1..%d destinations", kMaxMultiOut);
  for (int d = 0; d < n_out; ++d)
    if (outs[d] == nullptr || (reinterpret_cast<uintptr_t>(outs[d]) & 15))
      return fail(VR2_ERR_ARG, "vr2_tower_forward_multi: destination %d is null or not 16-byte aligned", d);
  return tower_forward_impl(t, plan, pixels, pixels_dtype, outs[0], VR2_BF16, nullptr, 0, workspace, workspace_bytes, stream_v,
                            outs, n_out);
}

extern "C" int vr2_tower_forward_scatter(vr2_tower* t, const vr2_plan* plan, const void* pixels, int32_t pixels_dtype,
                                         void* out, int32_t out_dtype, int64_t ld_out, const int32_t* dest_rows_dev,
                                         void* workspace, size_t workspace_bytes, void* stream_v) {
  if (dest_rows_dev == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_forward_scatter: null dest_rows");
  if (t != nullptr && (ld_out < t->cfg.out_hidden_size || ld_out % 8 != 0))
    return fail(VR2_ERR_ARG, "vr2_tower_forward_scatter: ld_out %lld must be >= out_hidden_size and a multiple of 8", (long long)ld_out);
  return tower_forward_impl(t, plan, pixels, pixels_dtype, out, out_dtype, dest_rows_dev, ld_out, workspace, workspace_bytes, stream_v);
}

extern "C" int vr2_tower_set_precision(vr2_tower* t, int32_t precision) {
  if (t == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_set_precision: null tower");
  if (precision != VR2_PRECISION_BF16 && precision != VR2_PRECISION_SPLIT && precision != VR2_PRECISION_F16)
    return fail(VR2_ERR_ARG, "vr2_tower_set_precision: unknown precision %d", precision);
  t->precision = precision;
  return VR2_OK;
}

extern "C" int32_t vr2_tower_get_precision(const vr2_tower* t) { return t == nullptr ? -1 : t->precision; }

extern "C" int vr2_tower_saturation_snapshot(vr2_tower* t, uint32_t* dst, void* stream_v) {
  if (t == nullptr || dst == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_saturation_snapshot: null argument");
  VR2_CUDA(cudaMemcpyAsync(dst, t->sat, VR2_SAT_SITES * sizeof(uint32_t), cudaMemcpyDefault, as_stream(stream_v)));
  return VR2_OK;
}

extern "C" int vr2_tower_saturation_counts(vr2_tower* t, uint64_t* counts_host, int32_t reset, void* stream_v) {
  if (t == nullptr || counts_host == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_saturation_counts: null argument");
  cudaStream_t s = as_stream(stream_v);
  unsigned int v[VR2_SAT_SITES] = {0, 0, 0, 0};
  VR2_CUDA(cudaMemcpyAsync(v, t->sat, sizeof(v), cudaMemcpyDeviceToHost, s));
  if (reset) VR2_CUDA(cudaMemsetAsync(t->sat, 0, sizeof(v), s));
  VR2_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < VR2_SAT_SITES; ++i) counts_host[i] = v[i];
  return VR2_OK;
}

extern "C" int vr2_tower_saturation_count(vr2_tower* t, uint64_t* count_host, int32_t reset, void* stream_v) {
  if (t == nullptr || count_host == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_saturation_count: null argument");
  uint64_t v[VR2_SAT_SITES];
  const int rc = vr2_tower_saturation_counts(t, v, reset, stream_v);
  if (rc != VR2_OK) return rc;
  *count_host = v[0] + v[1] + v[2] + v[3];
  return VR2_OK;
}

extern "C" int vr2_tower_set_f16_shift(vr2_tower* t, int32_t shift_xb, int32_t shift_h) {
  if (t == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_set_f16_shift: null tower");
  if (!t->finalized) return fail(VR2_ERR_STATE, "vr2_tower_set_f16_shift: weights not finalized");
  if (shift_xb < 0 || shift_xb > 14 || shift_h < 0 || shift_h > 14)
    return fail(VR2_ERR_ARG, "vr2_tower_set_f16_shift: shifts must be in [0, 14], got %d / %d", shift_xb, shift_h);
  int prev = -1;
  VR2_CUDA(cudaGetDevice(&prev));
  VR2_CUDA(cudaSetDevice(t->device));
  int rc = VR2_OK;
  auto upload = [&](Tensor* dst, const std::vector<float>& v) -> int {
    if (dst->ptr == nullptr) {
      void* p = nullptr;
      VR2_CUDA(cudaMalloc(&p, v.size() * sizeof(float)));
      t->allocs.push_back(p);
      dst->ptr = p;
    }
    VR2_CUDA(cudaMemcpy(dst->ptr, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));   // orders after prior work
    return VR2_OK;
  };
  VR2_CUDA(cudaDeviceSynchronize());   // no forward may be reading the copies that are about to change
  if (shift_xb != 0) {
    t->norm_s.resize(t->norm_host.size());
    const float c = ldexpf(1.f, -shift_xb);
    std::vector<float> tmp;
    for (size_t i = 0; i < t->norm_host.size() && rc == VR2_OK; ++i) {
      tmp = t->norm_host[i];
      for (float& v : tmp) v *= c;
      rc = upload(&t->norm_s[i], tmp);
    }
  }
  if (shift_h != 0 && rc == VR2_OK) {
    t->gub_s.resize(t->gub_host.size());
    const float c = ldexpf(1.f, -shift_h);
    std::vector<float> tmp;
    for (size_t l = 0; l < t->gub_host.size() && rc == VR2_OK; ++l) {
      tmp = t->gub_host[l];
      for (size_t j = 0; j < tmp.size(); ++j)
        if (j % 256 >= 128) tmp[j] *= c;      // the `up` half of every 256-column gate|up tile
      rc = upload(&t->gub_s[l], tmp);
    }
  }
  cudaSetDevice(prev);
  if (rc != VR2_OK) return rc;
  t->shift_xb = shift_xb;
  t->shift_h = shift_h;
  return VR2_OK;
}

extern "C" int vr2_tower_get_f16_shift(const vr2_tower* t, int32_t* shifts_host) {
  if (t == nullptr || shifts_host == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_get_f16_shift: null argument");
  shifts_host[0] = t->shift_xb;
  shifts_host[1] = t->shift_h;
  return VR2_OK;
}

extern "C" int vr2_tower_set_debug_stop(vr2_tower* t, int32_t n_stages) {
  if (t == nullptr || n_stages < 0) return fail(VR2_ERR_ARG, "vr2_tower_set_debug_stop: null tower or negative count");
  t->debug_stop = n_stages;
  return VR2_OK;
}

extern "C" int vr2_tower_workspace_layout(const vr2_tower* t, int64_t n_patches, int64_t* offsets, int64_t* h_ld) {
  if (t == nullptr || offsets == nullptr || h_ld == nullptr || n_patches <= 0)
    return fail(VR2_ERR_ARG, "vr2_tower_workspace_layout: null argument");
  char* base = reinterpret_cast<char*>(static_cast<uintptr_t>(1024));   // only differences of the carved pointers are used
  const Workspace w = carve_ws(t, n_patches, base);
  const void* p[7] = {w.x, w.xb, w.qkv, w.att, w.hbuf, w.ssq, w.pix};
  for (int i = 0; i < 7; ++i) offsets[i] = static_cast<const char*>(p[i]) - base;
  *h_ld = t->ipad;
  return VR2_OK;
}

extern "C" int vr2_tower_profile(vr2_tower* t, uint32_t kind_mask) {
  if (t == nullptr) return fail(VR2_ERR_ARG, "vr2_tower_profile: null tower");
  t->prof_mask = kind_mask;
  t->prof_used = 0;
  return VR2_OK;
}

extern "C" int vr2_tower_profile_read(vr2_tower* t, double* ms_by_kind, int64_t* launches_by_kind) {
  if (t == nullptr || ms_by_kind == nullptr || launches_by_kind == nullptr)
    return fail(VR2_ERR_ARG, "vr2_tower_profile_read: null argument");
  for (int k = 0; k < VR2_K_COUNT; ++k) { ms_by_kind[k] = 0.0; launches_by_kind[k] = 0; }
  for (size_t i = 0; i < t->prof_used; ++i) {
    VR2_CUDA(cudaEventSynchronize(t->prof_events[2 * i + 1]));
    float ms = 0.f;
    VR2_CUDA(cudaEventElapsedTime(&ms, t->prof_events[2 * i], t->prof_events[2 * i + 1]));
    ms_by_kind[t->prof_kind[i]] += ms;
    launches_by_kind[t->prof_kind[i]] += 1;
  }
  t->prof_used = 0;
  return VR2_OK;
}

// ----------------------------------------------------------------- single operators (test surface)
static int op_linear_impl(const void* a_dev, const void* a_lo_dev, const void* w_dev, const float* bias, void* out,
                          int32_t out_dtype, int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream_v,
                          bool a_f16 = false);

extern "C" int vr2_op_linear_f16(const void* a_dev, const void* w_dev, const float* bias, void* out, int32_t out_dtype,
                                 int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream_v) {
  return op_linear_impl(a_dev, nullptr, w_dev, bias, out, out_dtype, M, N, K, gelu, stream_v, true);
}

extern "C" int vr2_op_linear(const void* a_dev, const void* w_dev, const float* bias, void* out, int32_t out_dtype,
                             int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream_v) {
  return op_linear_impl(a_dev, nullptr, w_dev, bias, out, out_dtype, M, N, K, gelu, stream_v);
}

extern "C" int vr2_op_linear_split(const void* a_hi_dev, const void* a_lo_dev, const void* w_dev, const float* bias,
                                   void* out, int32_t out_dtype, int32_t M, int32_t N, int32_t K, int32_t gelu,
                                   void* stream_v) {
  if (a_lo_dev == nullptr) return fail(VR2_ERR_ARG, "vr2_op_linear_split: null a_lo");
  return op_linear_impl(a_hi_dev, a_lo_dev, w_dev, bias, out, out_dtype, M, N, K, gelu, stream_v);
}

static int op_linear_impl(const void* a_dev, const void* a_lo_dev, const void* w_dev, const float* bias, void* out,
                          int32_t out_dtype, int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream_v, bool a_f16) {
  if (a_dev == nullptr || w_dev == nullptr || out == nullptr) return fail(VR2_ERR_ARG, "vr2_op_linear: null pointer");
  if (out_dtype != VR2_F32 && out_dtype != VR2_BF16 && !(a_f16 && out_dtype == VR2_F16))
    return fail(VR2_ERR_ARG, "vr2_op_linear: out dtype");
  if (K % 8 != 0) return fail(VR2_ERR_ARG, "vr2_op_linear: K %% 8 != 0 (rows must be 16-byte aligned)");
  int dev = 0, sms = 148;
  VR2_CUDA(cudaGetDevice(&dev));
  VR2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GemmArgs a{};
  a.A = static_cast<const __nv_bfloat16*>(a_dev); a.lda = K;
  a.A_lo = static_cast<const __nv_bfloat16*>(a_lo_dev);
  a.W = static_cast<const __nv_bfloat16*>(w_dev); a.ldw = K;
  a.M = M; a.N = N; a.K = K;
  a.epilogue = EPI_STORE;
  a.flags = (gelu ? GEMM_GELU : 0) | (out_dtype == VR2_F32 ? GEMM_OUT_F32 : 0) | (a_f16 ? GEMM_A_F16 : 0) |
            (out_dtype == VR2_F16 ? GEMM_OUT_F16 : 0);
  a.bias = bias;
  a.out_bf16 = static_cast<__nv_bfloat16*>(out); a.out_f32 = static_cast<float*>(out); a.ld_out = N;
  const char* msg = "";
  cudaError_t e = gemm_launch(a, sms, as_stream(stream_v), &msg);
  if (e != cudaSuccess) return fail(e == cudaErrorInvalidValue ? VR2_ERR_ARG : VR2_ERR_CUDA, "vr2_op_linear: %s %s", cudaGetErrorString(e), msg);
  return VR2_OK;
}

extern "C" int vr2_op_rmsnorm(const float* x, const float* w, void* out, int64_t rows, int32_t dim, float eps,
                              void* stream_v) {
  if (x == nullptr || w == nullptr || out == nullptr) return fail(VR2_ERR_ARG, "vr2_op_rmsnorm: null pointer");
  if (dim % 4 != 0) return fail(VR2_ERR_ARG, "vr2_op_rmsnorm: dim %% 4 != 0");
  VR2_CUDA(rmsnorm_to_bf16(x, w, static_cast<__nv_bfloat16*>(out), nullptr, rows, dim, eps, as_stream(stream_v)));
  return VR2_OK;
}

static int op_attention_impl(const void* qkv, const void* qkv_lo, void* out, void* out_lo, const int32_t* cu,
                             int32_t n_seq, int32_t num_heads, void* stream_v, bool f16 = false);

extern "C" int vr2_op_attention_f16(const void* qkv, void* out, const int32_t* cu, int32_t n_seq, int32_t num_heads,
                                    void* stream_v) {
  return op_attention_impl(qkv, nullptr, out, nullptr, cu, n_seq, num_heads, stream_v, true);
}

extern "C" int vr2_op_attention(const void* qkv, void* out, const int32_t* cu, int32_t n_seq, int32_t num_heads,
                                void* stream_v) {
  return op_attention_impl(qkv, nullptr, out, nullptr, cu, n_seq, num_heads, stream_v);
}

extern "C" int vr2_op_attention_split(const void* qkv_hi, const void* qkv_lo, void* out_hi, void* out_lo,
                                      const int32_t* cu, int32_t n_seq, int32_t num_heads, void* stream_v) {
  if (qkv_lo == nullptr || out_lo == nullptr) return fail(VR2_ERR_ARG, "vr2_op_attention_split: null lo pointer");
  return op_attention_impl(qkv_hi, qkv_lo, out_hi, out_lo, cu, n_seq, num_heads, stream_v);
}

static int op_attention_impl(const void* qkv, const void* qkv_lo, void* out, void* out_lo, const int32_t* cu,
                             int32_t n_seq, int32_t num_heads, void* stream_v, bool f16) {
  if (qkv == nullptr || out == nullptr || cu == nullptr || n_seq <= 0) return fail(VR2_ERR_ARG, "vr2_op_attention: null pointer");
  for (int i = 0; i < n_seq; ++i)
    if (cu[i + 1] < cu[i]) return fail(VR2_ERR_ARG, "vr2_op_attention: cu_seqlens not monotone");
  cudaStream_t s = as_stream(stream_v);
  cudaError_t e = cudaSuccess;
  void* d = nullptr;
  {
    // per-row segment bounds (both kernels take them)
    const int n_rows = cu[n_seq] - cu[0];
    if (cu[0] != 0) return fail(VR2_ERR_ARG, "vr2_op_attention: cu_seqlens must start at 0");
    if (n_rows == 0) return VR2_OK;
    std::vector<int2> seg(n_rows);
    for (int i = 0; i < n_seq; ++i)
      for (int r = cu[i]; r < cu[i + 1]; ++r) seg[r] = make_int2(cu[i], cu[i + 1]);
    int dev = 0, sms = 148;
    VR2_CUDA(cudaGetDevice(&dev));
    VR2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    VR2_CUDA(cudaMalloc(&d, seg.size() * 8));
    e = cudaMemcpyAsync(d, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
      if (qkv_lo == nullptr)
        e = attention_tc_launch(static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out),
                                static_cast<const int2*>(d), n_rows, num_heads, 3 * num_heads * 80, num_heads * 80, sms, s, f16);
      else
        e = attention_split_launch(static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(qkv_lo),
                                   static_cast<__nv_bfloat16*>(out), static_cast<__nv_bfloat16*>(out_lo),
                                   static_cast<const int2*>(d), n_rows, num_heads, 3 * num_heads * 80, num_heads * 80, sms, s);
    }
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d);
  if (e != cudaSuccess) return fail(VR2_ERR_CUDA, "vr2_op_attention: %s", cudaGetErrorString(e));
  return VR2_OK;
}
