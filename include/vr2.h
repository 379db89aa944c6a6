/*
 * vr2.h — C ABI of libvr2.so: the B200 (sm_100a) implementation of the one hot
 * path Video-R2 spends its rollouts and evaluation in,
 *
 *   Qwen2_5_VisionTransformerPretrainedModel.forward(hidden_states, grid_thw)
 *   Qwen2_5_VLForConditionalGeneration.get_rope_index(input_ids, image_grid_thw,
 *        video_grid_thw, second_per_grid_ts, attention_mask)
 *
 * The reference (mbzuai-oryx/Video-R2) reaches both through Hugging Face
 * transformers==4.51.1 (requirements.txt:3); its call sites are
 *   train/src/train/monkey_patch_forward.py:366,374,392   self.visual(pixels, grid_thw=...)
 *   train/src/train/monkey_patch_forward.py:415-421       self.get_rope_index(...)
 *   train/src/trainer/grpo_trainer.py:725,783             (stock HF forward / generate)
 *   eval/lmms-eval/lmms_eval/models/qwen2_5_vl.py:326     (stock HF generate)
 * and it plugs replacements in by class-attribute assignment
 * (monkey_patch_forward.py:22-26).  video_r2_b200/patch.py does the same with
 * functions that call the entry points below through ctypes.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on
 * success and a negative vr2_status otherwise, with a thread-local message in
 * vr2_last_error().  Pointers named *_dev are CUDA device pointers of the
 * current device; everything is enqueued on `stream` (a cudaStream_t passed as
 * void*) and nothing synchronises unless stated.  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with VR2_ERR_CUDA.
 */
#ifndef VR2_H_
#define VR2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vr2_status {
  VR2_OK = 0,
  VR2_ERR_ARG = -1,    /* bad shape / dtype / null pointer (Python shim raises ValueError) */
  VR2_ERR_CUDA = -2,   /* CUDA runtime or driver error */
  VR2_ERR_STATE = -3,  /* weights missing, plan/tower mismatch */
} vr2_status;

typedef enum vr2_dtype {
  VR2_F32 = 0,
  VR2_BF16 = 1,
  VR2_I64 = 2,
  VR2_I32 = 3,
  VR2_U8 = 4, /* also torch.bool */
  VR2_F16 = 5, /* IEEE half: activations of VR2_PRECISION_F16 (single-operator test surface only) */
} vr2_dtype;

const char* vr2_last_error(void);
int vr2_abi_version(void);

/* ------------------------------------------------------------------------
 * Index plan: everything the tower derives from grid_thw alone.
 * Replaces rot_pos_emb (HF modeling_qwen2_5_vl.py:382-409), get_window_index
 * (:411-451), the unique_consecutive of cu_window_seqlens (:476) and the
 * cu_seqlens construction (:488-496).  All integer outputs are bit-exact.
 * ---------------------------------------------------------------------- */
typedef struct vr2_plan vr2_plan;

typedef struct vr2_plan_config {
  int32_t spatial_merge_size; /* 2 */
  int32_t window_size;        /* 112 (pixels) */
  int32_t patch_size;         /* 14 */
  int32_t head_dim;           /* 80: rotary table has head_dim/2 angles per patch */
  float rope_theta;           /* 10000 */
} vr2_plan_config;

/* grid_thw_host: [n_items,3] int64 on the HOST (t, h, w in patches). Builds the
 * device tables on `stream`; the plan owns them. */
int vr2_plan_create(const vr2_plan_config* cfg, const int64_t* grid_thw_host, int32_t n_items, void* stream,
                    vr2_plan** out);
void vr2_plan_destroy(vr2_plan* plan);
int64_t vr2_plan_num_patches(const vr2_plan* plan);     /* N = sum t*h*w */
int64_t vr2_plan_num_tokens(const vr2_plan* plan);      /* N / merge^2 */
int64_t vr2_plan_num_windows(const vr2_plan* plan);     /* len(cu_window_seqlens) - 1 after dedupe */
int64_t vr2_plan_num_slices(const vr2_plan* plan);      /* len(cu_seqlens) - 1 = sum t */
/* Copies the integer tables into caller buffers (device). Any pointer may be NULL.
 *   window_index_dev  [tokens]      int64   == HF get_window_index()[0]
 *   reverse_index_dev [tokens]      int64   == torch.argsort(window_index)
 *   cu_window_dev     [windows+1]   int32   == unique_consecutive(cu_window_seqlens)
 *   cu_seqlens_dev    [slices+1]    int32   == F.pad(repeat_interleave(h*w, t).cumsum(0), (1,0))
 *   pos_hw_dev        [patches,2]   int32   == rot_pos_emb's pos_ids (h, w), original patch order */
int vr2_plan_export(const vr2_plan* plan, int64_t* window_index_dev, int64_t* reverse_index_dev,
                    int32_t* cu_window_dev, int32_t* cu_seqlens_dev, int32_t* pos_hw_dev, void* stream);
/* Rotary angle tables in WINDOW order, fp32 [patches, head_dim/2] each. */
int vr2_plan_export_rotary(const vr2_plan* plan, float* cos_dev, float* sin_dev, void* stream);

/* ------------------------------------------------------------------------
 * Bit-exact row gathers (the window permute HF :478-484 and its inverse
 * :512-513): dst[i, :] = src[index[i], :], rows of row_bytes (multiple of 16).
 * ---------------------------------------------------------------------- */
int vr2_gather_rows(const void* src_dev, void* dst_dev, const int64_t* index_dev, int64_t n_rows_out,
                    int64_t n_rows_in, int64_t row_bytes, void* stream);

/* ------------------------------------------------------------------------
 * The vision tower.
 * ---------------------------------------------------------------------- */
typedef struct vr2_tower vr2_tower;

typedef struct vr2_tower_config {
  int32_t depth;              /* 32 */
  int32_t hidden_size;        /* 1280 */
  int32_t num_heads;          /* 16 */
  int32_t intermediate_size;  /* 3420 */
  int32_t out_hidden_size;    /* 3584 */
  int32_t patch_size;         /* 14 */
  int32_t temporal_patch_size;/* 2 */
  int32_t in_channels;        /* 3 */
  int32_t spatial_merge_size; /* 2 */
  int32_t window_size;        /* 112 */
  int32_t n_fullatt;          /* number of valid entries in fullatt_block_indexes */
  int32_t fullatt_block_indexes[8]; /* 7,15,23,31 */
  float rms_norm_eps;         /* 1e-6 */
  float rope_theta;           /* 10000 */
} vr2_tower_config;

int vr2_tower_create(const vr2_tower_config* cfg, int32_t cuda_device, vr2_tower** out);
/* Operand precision of the dense contractions (GEMMs and attention); accumulation, residual stream, norms,
 * rotary and softmax are fp32 in both.
 *   VR2_PRECISION_BF16  (default): activations rounded to bf16 once per GEMM input — the arithmetic HF's own
 *                        bf16 path uses (modeling_qwen2_5_vl.py under autocast / .to(bfloat16)).
 *   VR2_PRECISION_SPLIT: every activation operand is carried as hi + lo bf16 (lo = rounding residual) and
 *                        multiplied twice (3 products inside attention): ~16 mantissa bits, 2x the MMA work.
 *                        Meets max-abs 2e-2 against the fp32 reference on a 32-block tower.
 *   VR2_PRECISION_F16:   activations (the operand of every contraction that is not a weight, the attention
 *                        probabilities, q/k/v) are rounded to IEEE fp16 (11 significand bits) instead of bf16 (8);
 *                        the GEMMs read fp16 copies of the bf16 weights (exact inside fp16's normal range; +1.35 GB,
 *                        made on first use), the patch embedding still multiplies bf16 pixels by bf16 weights.  The
 *                        work and the speed are those of VR2_PRECISION_BF16 with ~8x less operand rounding: also
 *                        meets 2e-2.  Conversions saturate at +-65504 (fp16's range) and every epilogue that
 *                        converts COUNTS the values it clamped (NaN included) in a device counter:
 *                        vr2_tower_saturation_count.  A forward whose count is non-zero must be repeated in SPLIT
 *                        (the Python host's default "auto" precision does exactly that).
 * Changes vr2_tower_workspace_bytes(); call before sizing the workspace. */
typedef enum vr2_precision { VR2_PRECISION_BF16 = 0, VR2_PRECISION_SPLIT = 1, VR2_PRECISION_F16 = 2 } vr2_precision;
int vr2_tower_set_precision(vr2_tower* tower, int32_t precision);
int32_t vr2_tower_get_precision(const vr2_tower* tower);
/* Range guard of VR2_PRECISION_F16: *count_host = number of (thread, tile) conversions since the last reset that met a
 * value outside fp16's finite range, over every forward enqueued on `stream` so far (the call synchronises `stream`).
 * 0 means no activation was clamped: the result is what unbounded-range fp16 arithmetic would give.  reset != 0
 * zeroes the counter (stream-ordered).  Always 0 in the other precisions. */
int vr2_tower_saturation_count(vr2_tower* tower, uint64_t* count_host, int32_t reset, void* stream);
/* The same, per conversion site: counts_host[VR2_SAT_SITES] =
 *   [0] xb, the residual-stream operand x * w_norm written by the patch-embed / proj / down epilogues,
 *   [1] q / k / v,   [2] h = silu(gate) * up,   [3] the merger's ln_q output and fc1 hidden. */
#define VR2_SAT_SITES 4
int vr2_tower_saturation_counts(vr2_tower* tower, uint64_t* counts_host, int32_t reset, void* stream);
/* The VR2_SAT_SITES counters copied (uint32 each), stream-ordered and without synchronising, to a 16-byte destination
 * in pinned host or device memory: a pipelined caller snapshots them behind every forward and compares neighbours
 * once the batch is complete. */
int vr2_tower_saturation_snapshot(vr2_tower* tower, uint32_t* dst, void* stream);
/* Range shifts of VR2_PRECISION_F16, for checkpoints whose residual stream or SwiGLU products exceed fp16's range
 * (massive activations): xb is stored times 2^-shift_xb and h times 2^-shift_h, and the consuming GEMMs' epilogues
 * multiply the fp32 accumulator back (powers of two: exact; values below 2^(shift - 14) lose relative precision,
 * far below the fp16 rounding of the O(1) entries).  Same instructions, same speed; the default 0 / 0 is bit-identical
 * to the unshifted arithmetic.  Synchronises the device (the scaled norm-weight / bias copies are rebuilt). */
int vr2_tower_set_f16_shift(vr2_tower* tower, int32_t shift_xb, int32_t shift_h);
int vr2_tower_get_f16_shift(const vr2_tower* tower, int32_t* shifts_host /* [2] */);
void vr2_tower_destroy(vr2_tower* tower);
/* One call per entry of the HF module's state_dict (names relative to the
 * vision module: "patch_embed.proj.weight", "blocks.0.norm1.weight",
 * "blocks.0.attn.qkv.weight", ..., "merger.mlp.2.bias").  The tensor is read on
 * `stream`, converted and repacked into library-owned bf16/fp32 storage; the
 * caller's buffer is not retained.  dtype: VR2_F32 or VR2_BF16. */
int vr2_tower_set_weight(vr2_tower* tower, const char* name, const void* data_dev, int32_t dtype,
                         const int64_t* shape, int32_t ndim, void* stream);
/* Fails with VR2_ERR_STATE and lists what is missing unless every tensor was set. */
int vr2_tower_finalize(vr2_tower* tower);
size_t vr2_tower_workspace_bytes(const vr2_tower* tower, int64_t n_patches);
/* pixels_dev: [N, 1176] (dtype VR2_F32 or VR2_BF16), rows ordered as the HF
 * processor emits them; out_dev: [N/4, out_hidden] (VR2_BF16 or VR2_F32) in the
 * ORIGINAL token order (inverse window permute applied), i.e. exactly what
 * transformers 4.51.1 returns from visual.forward. workspace_dev: at least
 * vr2_tower_workspace_bytes() bytes, 1024-byte aligned. */
int vr2_tower_forward(vr2_tower* tower, const vr2_plan* plan, const void* pixels_dev, int32_t pixels_dtype,
                      void* out_dev, int32_t out_dtype, void* workspace_dev, size_t workspace_bytes, void* stream);
/* Same forward, with the caller's masked_scatter fused into the merger's last store: merged token k (original order)
 * is written to out_dev[dest_rows_dev[k], 0:out_hidden] of a buffer with row pitch ld_out elements — e.g.
 * inputs_embeds viewed as [B*S, hidden] with dest_rows = the flat positions of <|video_pad|> (vr2_token_positions).
 * Replaces Video-R2 train/src/train/monkey_patch_forward.py:385-406 (masked_scatter of video_embeds) together with
 * the tower call at :392; rows not named in dest_rows are left untouched, tokens whose dest_rows entry is negative
 * are not stored. */
int vr2_tower_forward_scatter(vr2_tower* tower, const vr2_plan* plan, const void* pixels_dev, int32_t pixels_dtype,
                              void* out_dev, int32_t out_dtype, int64_t ld_out, const int32_t* dest_rows_dev,
                              void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same forward for the SHARDED multi-GPU mode: the bf16 result [n_tokens, out_hidden] is stored to
 * n_out (1..8) buffers at once — outs_host[d] is a device pointer to row 0 of this call's tokens in destination d
 * (row pitch out_hidden), 16-byte aligned; typically this rank's copy of the gathered embeddings and the peer GPUs'
 * copies of it, mapped into this process (CUDA IPC / symmetric memory), each already offset by this rank's first
 * token.  The merger's last GEMM stages every 128 x 256 output tile in shared memory and sends each row to every
 * destination with one 512-byte bulk copy (cp.async.bulk over NVLink for peer destinations): the all-gather of the
 * merged embeddings overlaps the GEMM tile by tile and needs no collective behind it — only a barrier across the
 * ranks before the result is read (the kernel ends with a system-scope fence after its copies complete).
 * No reference counterpart: Video-R2 encodes the same clip on every rank (grpo_trainer.py:664-679). */
int vr2_tower_forward_multi(vr2_tower* tower, const vr2_plan* plan, const void* pixels_dev, int32_t pixels_dtype,
                            void* const* outs_host, int32_t n_out, void* workspace_dev, size_t workspace_bytes,
                            void* stream);

/* Test hooks (stage-by-stage parity, tests/test_gpu_stages.py).  vr2_tower_set_debug_stop(n): every following forward
 * returns after n stages — 1 = patch embed, then qkv, attention, proj, gate|up, down of block 0, and so on — leaving the
 * caller's workspace as that stage wrote it (0 = off, the default).  vr2_tower_workspace_layout: byte offsets of the
 * workspace buffers for a forward of n_patches rows, offsets[0..6] = x (fp32 [N, hidden], residual stream), xb (16-bit
 * [N, hidden], x times the next norm weight), qkv (16-bit [N, 3 hidden]), att (16-bit [N, hidden]), h (16-bit [N, *h_ld],
 * SwiGLU output, columns >= intermediate_size are padding), ssq (fp32 [parts][N] partial sums of x^2, parts =
 * hidden / 256 from the CTA-pair kernel), pix (pixel staging).  All in WINDOW order (vr2_plan_export). */
int vr2_tower_set_debug_stop(vr2_tower* tower, int32_t n_stages);
int vr2_tower_workspace_layout(const vr2_tower* tower, int64_t n_patches, int64_t* offsets, int64_t* h_ld);

/* Per-kernel-kind device timing.  vr2_tower_profile(mask) selects kinds (bit k = vr2_kernel_kind k; 0 turns it
 * off) and resets the counters; while selected, vr2_tower_forward brackets every launch of that kind with a
 * cudaEvent pair on the caller's stream.  vr2_tower_profile_read synchronises on those events and returns the
 * summed milliseconds and launch counts per kind (arrays of VR2_K_COUNT), then resets. */
typedef enum vr2_kernel_kind {
  VR2_K_CAST = 0,
  VR2_K_PATCH_EMBED = 1,
  VR2_K_QKV = 2,
  VR2_K_ATTN_WINDOW = 3,
  VR2_K_ATTN_FULL = 4,
  VR2_K_PROJ = 5,
  VR2_K_GATE_UP = 6,
  VR2_K_DOWN = 7,
  VR2_K_MERGER_NORM = 8,
  VR2_K_FC1 = 9,
  VR2_K_FC2 = 10,
  VR2_K_COUNT = 11
} vr2_kernel_kind;
int vr2_tower_profile(vr2_tower* tower, uint32_t kind_mask);
int vr2_tower_profile_read(vr2_tower* tower, double* ms_by_kind, int64_t* launches_by_kind);
/* Number of kernels vr2_tower_forward enqueues for this plan (for gpu_launches). */
int64_t vr2_tower_launch_count(const vr2_tower* tower, const vr2_plan* plan, int32_t pixels_dtype);

/* ------------------------------------------------------------------------
 * get_rope_index, transformers 4.51.1 semantics.
 * ---------------------------------------------------------------------- */
typedef struct vr2_rope_config {
  int32_t spatial_merge_size;   /* 2 */
  int64_t image_token_id;       /* 151655 */
  int64_t video_token_id;       /* 151656 */
  int64_t vision_start_token_id;/* 151652 */
  float tokens_per_second;      /* 2 */
  /* how second_per_grid_t enters the temporal index:
   *   0: t = trunc(fp32(fp32(k * spg) * tps))      (float multiply, then .long())
   *   1: t = k * trunc(spg) * tps                  (as_tensor(..., dtype=int64) first) */
  int32_t spg_int_cast;
} vr2_rope_config;

/* input_ids_dev [B,S] int64; mask_dev [B,S] of mask_dtype (I64/I32/U8) or NULL;
 * image_thw_dev [n_img,3] / video_thw_dev [n_vid,3] int64 or NULL;
 * second_per_grid_ts_dev [n_vid] fp32 or NULL;
 * scratch_dev: vr2_rope_scratch_bytes(B, S) bytes;
 * position_ids_dev [3,B,S] int64; deltas_dev [B] int64.
 * With both grids NULL this is the text-only branch.
 * A row that names more images / videos than the grids describe (or names one while that grid pointer is NULL)
 * makes Hugging Face raise; the kernel never reads through a NULL or past-the-end grid pointer: the item is not
 * placed, the rest of the row is numbered as text and a status word is set in the scratch buffer.
 * vr2_rope_index_status copies it to the host (synchronises `stream`): 0 = every item placed, bit 0 = image grid
 * missing / too short, bit 1 = video grid missing / too short. */
size_t vr2_rope_scratch_bytes(int32_t B, int32_t S);
int vr2_rope_index(const vr2_rope_config* cfg, const int64_t* input_ids_dev, const void* mask_dev, int32_t mask_dtype,
                   int32_t B, int32_t S, const int64_t* image_thw_dev, int32_t n_img, const int64_t* video_thw_dev,
                   int32_t n_vid, const float* second_per_grid_ts_dev, void* scratch_dev, int64_t* position_ids_dev,
                   int64_t* deltas_dev, void* stream);
int vr2_rope_index_status(const void* scratch_dev, int32_t B, int32_t S, int32_t* status_host, void* stream);

/* ------------------------------------------------------------------------
 * Placing visual tokens for the decoder (monkey_patch_forward.py:372-406).
 * ---------------------------------------------------------------------- */
/* positions_dev[k] = flat index of the k-th element of input_ids_dev[0:n] equal to token_id (k < capacity), i.e. the
 * order in which `mask = input_ids == token_id; x.masked_scatter(mask[..., None], embeds)` consumes the rows of
 * embeds; entries past the number of matches are -1 (consumers skip them).  The number of matches (may exceed
 * capacity; HF raises "features and tokens do not match" when it differs from the number of features) is returned
 * in *count_host, which synchronises `stream` — or pass count_host = NULL to stay asynchronous and fetch it later
 * with vr2_token_positions_count(scratch_dev, n, ...), e.g. after the tower has been enqueued.
 * scratch_dev: vr2_token_positions_scratch_bytes(n). */
size_t vr2_token_positions_scratch_bytes(int64_t n);
int vr2_token_positions(const int64_t* input_ids_dev, int64_t n, int64_t token_id, int32_t* positions_dev,
                        int64_t capacity, void* scratch_dev, int64_t* count_host, void* stream);
int vr2_token_positions_count(const void* scratch_dev, int64_t n, int64_t* count_host, void* stream);
/* dst[dest_rows[k], 0:row_elems] = convert(src[k, 0:row_elems]) for k < n_rows; src rows contiguous, dst row pitch
 * ld_dst elements; dtypes fp32 / bf16; row_elems and ld_dst multiples of 8.  The scatter for cached embeddings. */
int vr2_scatter_rows(const void* src_dev, int32_t src_dtype, void* dst_dev, int32_t dst_dtype,
                     const int32_t* dest_rows_dev, int64_t n_rows, int64_t row_elems, int64_t ld_dst, void* stream);

/* ------------------------------------------------------------------------
 * Frames -> tower input rows (after the resize): rescale + normalise + patchify of
 * transformers Qwen2VLVideoProcessor._preprocess and the caller's cast to the tower dtype.
 * frames_dev: [T, C, H, W] uint8 or fp32 (H, W multiples of patch*merge); mean_host / std_host: [C] fp32 ALREADY
 * fused with the rescale factor as HF does (mean / rescale_factor, std / rescale_factor in fp32);
 * out_dev: [ceil(T/Tp) * (H/P) * (W/P), C*Tp*P*P] fp32 or bf16, rows in the processor's order.  T not a multiple of
 * temporal_patch_size repeats the last frame.  Bit-exact with the processor (fp32) and with its output cast to bf16. */
int vr2_patchify(const void* frames_dev, int32_t frames_dtype, int32_t T, int32_t C, int32_t H, int32_t W,
                 int32_t patch_size, int32_t temporal_patch_size, int32_t merge_size, const float* mean_host,
                 const float* std_host, void* out_dev, int32_t out_dtype, void* stream);

/* ------------------------------------------------------------------------
 * Frame resize (before patchify): replaces the host-side
 * `transforms.functional.resize(video, [h, w], BICUBIC, antialias=True)` of Video-R2's fetch_video
 * (train/qwen_vl_utils/vision_process.py:802-807; same under demo/ and eval/).  frames_dev: uint8, element
 * (t, c, y, x) at byte offset t*stride_t + c*stride_c + y*stride_h + x*stride_w (so [T,C,H,W] and the decoders'
 * [T,H,W,C] are both accepted without a copy); out_dev: contiguous [T, C, out_h, out_w], uint8 (torchvision's result)
 * or fp32 (the same values, as `.float()` returns them).  workspace_dev: vr2_resize_workspace_bytes() bytes, holds
 * the two filter tables of this call.  T*C <= 65535 per call.  Byte-identical to torchvision 0.26 / torch 2.11 on
 * the CPU (oracle/resize.py). */
int vr2_resize_workspace_bytes(int32_t H, int32_t W, int32_t out_h, int32_t out_w, size_t* bytes);
int vr2_resize_frames(const void* frames_dev, int32_t T, int32_t C, int32_t H, int32_t W, int64_t stride_t,
                      int64_t stride_c, int64_t stride_h, int64_t stride_w, int32_t out_h, int32_t out_w, void* out_dev,
                      int32_t out_dtype, void* workspace_dev, void* stream);

/* ------------------------------------------------------------------------
 * Decoder-side M-RoPE application: the consumer of vr2_rope_index's position ids.
 * Replaces Qwen2_5_VLRotaryEmbedding.forward + apply_multimodal_rotary_pos_emb (transformers
 * modeling_qwen2_5_vl.py; Liger's liger_multimodal_rotary_pos_emb in train_sft.py:83-84) with one kernel that forms
 * cos/sin from position_ids on the fly.  q_dev [B, Hq, S, D], k_dev [B, Hk, S, D] (contiguous, fp32 or bf16; the
 * outputs may alias the inputs), position_ids_dev [3, B, S] int64, inv_freq_dev [D/2] fp32 (the module's buffer),
 * mrope_section_host [3] (sums to D/2; [16, 24, 24] for Qwen2.5-VL-7B), attention_scaling (1.0 for default rope).
 * bf16: every product and the sum are rounded to bf16 as in HF's bf16 forward. */
int vr2_mrope_apply(const void* q_dev, const void* k_dev, void* q_out_dev, void* k_out_dev, int32_t dtype, int32_t B,
                    int32_t S, int32_t Hq, int32_t Hk, int32_t D, const int64_t* position_ids_dev,
                    const float* inv_freq_dev, const int32_t* mrope_section_host, float attention_scaling, void* stream);

/* The same rotation FUSED into the decoder's q | k | v projection (the consumer side of
 * vr2_rope_index inside the LLM: Qwen2_5_VLAttention.forward, modeling_qwen2_5_vl.py:627-669 — q_proj / k_proj / v_proj
 * (Linear with bias), `.view(B, S, heads, 128)`, apply_multimodal_rotary_pos_emb on q and k):
 *
 *   vr2_mrope_table        once per decoder forward: cos / sin of the 64 rotary angles of every token, fp32
 *                          [B*S, head_dim/2] each, mrope_section axis choice applied (what Qwen2_5_VLRotaryEmbedding.
 *                          forward + the split / cat of apply_multimodal_rotary_pos_emb produce, never cast to bf16);
 *   vr2_decoder_qkv_mrope  per layer: out[M, N] = bf16(rot(hidden[M, K] · wqkv[N, K]^T + bias)) on the CTA-pair tcgen05
 *                          GEMM; wqkv / bias are the q, k, v projections concatenated along N (N = (Hq + 2 Hkv) * 128);
 *                          columns < rot_cols (= (Hq + Hkv) * 128) are rotated head by head with pairs (d, d + 64), in
 *                          fp32 on the accumulator, rounded to bf16 once.  q = out[:, :Hq*128].view(B, S, Hq, 128)
 *                          .transpose(1, 2) etc. are the (strided) tensors HF has at that point.
 * hidden / wqkv bf16, 16-byte aligned; N % 256 == 0; head_dim 128.  More accurate than the reference's own bf16
 * pipeline (which rounds after the Linear, after each product and after the sum): compared against fp32 in the tests. */
int vr2_mrope_table(const int64_t* position_ids_dev, int32_t B, int32_t S, const float* inv_freq_dev, int32_t half_dim,
                    const int32_t* mrope_section_host, float attention_scaling, float* cos_dev, float* sin_dev, void* stream);
int vr2_decoder_qkv_mrope(const void* hidden_dev, const void* wqkv_dev, const float* bias_dev, int64_t M, int32_t N, int32_t K,
                          int32_t head_dim, int32_t rot_cols, const float* cos_dev, const float* sin_dev, void* out_dev,
                          void* stream);

/* ------------------------------------------------------------------------
 * LM-head per-token log-probabilities without materialising the logits.
 * out[m] = log_softmax(hidden[m] @ W^T / temperature)[target[m]], fp32.  Replaces, in Video-R2's GRPO
 * (train/src/trainer/grpo_trainer.py:723-741), lm_head -> logits / temperature -> selective_log_softmax, which writes
 * and re-reads a [tokens, vocab] bf16 tensor.  hidden_dev [M, K] bf16, weight_dev [V, K] bf16 (lm_head.weight), target_dev
 * [M] int64 (a target outside [0, V) yields NaN), workspace_dev: vr2_lm_logprob_workspace_bytes(M, V), 16-byte aligned.
 * The contraction runs on the CTA-pair tcgen05 GEMM with fp32 accumulation; scores are not rounded to bf16. */
size_t vr2_lm_logprob_workspace_bytes(int64_t M, int32_t V);
int vr2_lm_logprob(const void* hidden_dev, const void* weight_dev, const int64_t* target_dev, int64_t M, int32_t V,
                   int32_t K, float temperature, float* out_logprob_dev, void* workspace_dev, size_t workspace_bytes,
                   void* stream);

/* ------------------------------------------------------------------------
 * Single operators, exported so tests can pin each kernel against the oracle.
 * ---------------------------------------------------------------------- */
/* C[M,N] = A[M,K] W[N,K]^T (+bias) (GELU optional); A, W bf16; out bf16 or fp32. */
int vr2_op_linear(const void* a_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t out_dtype,
                  int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream);
/* Split-operand variant: C = (A_hi + A_lo) W^T, A_lo bf16 with A's layout. */
int vr2_op_linear_split(const void* a_hi_dev, const void* a_lo_dev, const void* w_dev, const float* bias_dev,
                        void* out_dev, int32_t out_dtype, int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream);
/* fp16 variant (VR2_PRECISION_F16): A and W fp16 (tcgen05 kind::f16 rejects mixed fp16 x bf16 operands — the tower
 * keeps fp16 copies of its bf16 weights for this mode); out fp32, bf16 or fp16 (saturating at +-65504). */
int vr2_op_linear_f16(const void* a_f16_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t out_dtype,
                      int32_t M, int32_t N, int32_t K, int32_t gelu, void* stream);
/* out = bf16(x * rsqrt(mean(x^2)+eps) * w), x fp32 [rows, dim]. */
int vr2_op_rmsnorm(const float* x_dev, const float* w_dev, void* out_bf16_dev, int64_t rows, int32_t dim, float eps,
                   void* stream);
/* Varlen attention: qkv bf16 [N, 3*heads*80] (rotary already applied), cu_seqlens_host int32 [n_seq+1] (HOST),
 * out bf16 [N, heads*80]. Synchronises the stream (test helper). */
int vr2_op_attention(const void* qkv_dev, void* out_dev, const int32_t* cu_seqlens_host, int32_t n_seq,
                     int32_t num_heads, void* stream);

/* fp16 variant (VR2_PRECISION_F16): qkv and out hold fp16. */
int vr2_op_attention_f16(const void* qkv_f16_dev, void* out_f16_dev, const int32_t* cu_seqlens_host, int32_t n_seq,
                         int32_t num_heads, void* stream);

/* Split-operand variant: q/k/v and the output are hi + lo bf16 pairs of the same layout. */
int vr2_op_attention_split(const void* qkv_hi_dev, const void* qkv_lo_dev, void* out_hi_dev, void* out_lo_dev,
                           const int32_t* cu_seqlens_host, int32_t n_seq, int32_t num_heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VR2_H_ */
