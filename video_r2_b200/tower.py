"""Host-side mirror of `Qwen2_5_VisionTransformerPretrainedModel` (transformers
modeling_qwen2_5_vl.py:345-518) over the C ABI in include/vr2.h.

PyTorch is used for device memory, streams and nothing else: every kernel is in
libvr2.so.  `VisionTower.forward(hidden_states, grid_thw)` keeps the 4.51.1
contract Video-R2 relies on (train/src/train/monkey_patch_forward.py:392): a
plain tensor `[N/4, out_hidden]`, same dtype as the input, original token order.
"""
from __future__ import annotations

import collections
import ctypes as C
import dataclasses
import os
import warnings
import weakref

import torch

from . import _lib
from ._lib import VR2_BF16, VR2_F32, check, lib


@dataclasses.dataclass(frozen=True)
class VisionConfig:
    """Qwen2.5-VL-7B vision_config.  Set explicitly: the HF class defaults differ
    (hidden_size 3584, tokens_per_second 4)."""
    depth: int = 32
    hidden_size: int = 1280
    num_heads: int = 16
    intermediate_size: int = 3420
    out_hidden_size: int = 3584
    patch_size: int = 14
    temporal_patch_size: int = 2
    in_channels: int = 3
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: tuple = (7, 15, 23, 31)
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    tokens_per_second: float = 2

    @classmethod
    def from_hf(cls, hf_cfg):
        get = lambda k, d: getattr(hf_cfg, k, d)
        return cls(
            depth=hf_cfg.depth, hidden_size=hf_cfg.hidden_size, num_heads=hf_cfg.num_heads,
            intermediate_size=hf_cfg.intermediate_size, out_hidden_size=hf_cfg.out_hidden_size,
            patch_size=hf_cfg.patch_size, temporal_patch_size=get("temporal_patch_size", 2),
            in_channels=get("in_channels", get("in_chans", 3)), spatial_merge_size=hf_cfg.spatial_merge_size,
            window_size=hf_cfg.window_size, fullatt_block_indexes=tuple(hf_cfg.fullatt_block_indexes),
            rms_norm_eps=1e-6, rope_theta=10000.0, tokens_per_second=get("tokens_per_second", 2),
        )

    @property
    def patch_dim(self):
        return self.in_channels * self.temporal_patch_size * self.patch_size * self.patch_size

    def weight_shapes(self) -> dict:
        """name -> shape of every tensor of the HF vision module's state_dict (modeling_qwen2_5_vl.py:91-146, :207-321),
        in the order `load_state_dict` accepts them: what `vr2_tower_set_weight` must be fed before finalize."""
        h, i, oh = self.hidden_size, self.intermediate_size, self.out_hidden_size
        md = h * self.spatial_merge_size ** 2
        shapes = {"patch_embed.proj.weight": (h, self.in_channels, self.temporal_patch_size, self.patch_size, self.patch_size)}
        for l in range(self.depth):
            p = f"blocks.{l}."
            shapes.update({
                p + "norm1.weight": (h,), p + "norm2.weight": (h,),
                p + "attn.qkv.weight": (3 * h, h), p + "attn.qkv.bias": (3 * h,),
                p + "attn.proj.weight": (h, h), p + "attn.proj.bias": (h,),
                p + "mlp.gate_proj.weight": (i, h), p + "mlp.gate_proj.bias": (i,),
                p + "mlp.up_proj.weight": (i, h), p + "mlp.up_proj.bias": (i,),
                p + "mlp.down_proj.weight": (h, i), p + "mlp.down_proj.bias": (h,),
            })
        shapes.update({"merger.ln_q.weight": (h,), "merger.mlp.0.weight": (md, md), "merger.mlp.0.bias": (md,),
                       "merger.mlp.2.weight": (oh, md), "merger.mlp.2.bias": (oh,)})
        return shapes

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return VR2_F32
    if dt == torch.bfloat16:
        return VR2_BF16
    raise ValueError(f"unsupported dtype {dt}: the tower takes float32 or bfloat16")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class Plan:
    """All index tables derived from grid_thw (window index, inverse, cu_seqlens,
    rotary tables, attention tile schedule), built on the device once per distinct grid."""

    def __init__(self, cfg: VisionConfig, grid_thw, device):
        if isinstance(grid_thw, torch.Tensor):
            grid = grid_thw.detach().to("cpu", torch.int64).reshape(-1, 3).tolist()  # one small D2H, as HF's .tolist()
        else:
            grid = [[int(v) for v in row] for row in grid_thw]
        if len(grid) == 0:
            raise ValueError("grid_thw is empty")
        self.grid = tuple(tuple(r) for r in grid)
        self.device = torch.device(device)
        flat = (C.c_int64 * (3 * len(grid)))(*[v for row in grid for v in row])
        pc = _lib.PlanConfig(cfg.spatial_merge_size, cfg.window_size, cfg.patch_size, cfg.head_dim, cfg.rope_theta)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().vr2_plan_create(C.byref(pc), flat, len(grid), _stream_ptr(self.device), C.byref(handle)),
                  "vr2_plan_create")
        self._h = handle
        self.half_dim = cfg.head_dim // 2
        self.num_patches = lib().vr2_plan_num_patches(handle)
        self.num_tokens = lib().vr2_plan_num_tokens(handle)
        self.num_windows = lib().vr2_plan_num_windows(handle)
        self.num_slices = lib().vr2_plan_num_slices(handle)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().vr2_plan_destroy(h)
            except Exception:
                pass

    def export(self):
        """-> dict of torch tensors (device): window_index, reverse_index, cu_window_seqlens,
        cu_seqlens, pos_hw — the bit-exact counterparts of HF's helpers."""
        d = self.device
        out = {
            "window_index": torch.empty(self.num_tokens, dtype=torch.int64, device=d),
            "reverse_index": torch.empty(self.num_tokens, dtype=torch.int64, device=d),
            "cu_window_seqlens": torch.empty(self.num_windows + 1, dtype=torch.int32, device=d),
            "cu_seqlens": torch.empty(self.num_slices + 1, dtype=torch.int32, device=d),
            "pos_hw": torch.empty(self.num_patches, 2, dtype=torch.int32, device=d),
        }
        with torch.cuda.device(d):
            check(lib().vr2_plan_export(self._h, out["window_index"].data_ptr(), out["reverse_index"].data_ptr(),
                                        out["cu_window_seqlens"].data_ptr(), out["cu_seqlens"].data_ptr(),
                                        out["pos_hw"].data_ptr(), _stream_ptr(d)), "vr2_plan_export")
        return out

    def export_rotary(self):
        d = self.device
        cos = torch.empty(self.num_patches, self.half_dim, dtype=torch.float32, device=d)
        sin = torch.empty_like(cos)
        with torch.cuda.device(d):
            check(lib().vr2_plan_export_rotary(self._h, cos.data_ptr(), sin.data_ptr(), _stream_ptr(d)),
                  "vr2_plan_export_rotary")
        return cos, sin


def gather_rows(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """dst[i] = src[index[i]] for a 2-D contiguous src — the window permute / inverse as a
    bit-exact HBM gather (HF :478-484, :512-513)."""
    if src.dim() != 2 or not src.is_contiguous() or not src.is_cuda:
        raise ValueError("gather_rows: src must be a contiguous 2-D CUDA tensor")
    index = index.to(device=src.device, dtype=torch.int64).contiguous()
    row_bytes = src.shape[1] * src.element_size()
    dst = torch.empty((index.numel(), src.shape[1]), dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().vr2_gather_rows(src.data_ptr(), dst.data_ptr(), index.data_ptr(), index.numel(), src.shape[0],
                                    row_bytes, _stream_ptr(src.device)), "vr2_gather_rows")
    return dst


class VisionTower:
    """Owns a vr2_tower handle and its repacked weights.

    state_dict: name -> tensor with the HF vision-module names
    ("patch_embed.proj.weight", "blocks.0.attn.qkv.weight", ...), fp32 or bf16, any device."""

    def __init__(self, cfg: VisionConfig, state_dict, device="cuda:0", plan_cache: int = 16, precision=None):
        """precision:
          "auto"  (default) fp16 activations with the library's saturation guard: every forward is followed by a read of
                  the guard counters (one 16-byte D2H + stream synchronisation, where HF's own forward has its
                  `.tolist()` syncs).  If an activation left fp16's range the call is repeated: first with the range
                  SHIFTED at the site that clamped (the residual-stream operand and / or the SwiGLU product stored times
                  2^-4, 2^-8, 2^-12 — exact, same speed, remembered for later calls: a checkpoint with massive
                  activations pays for this once), and with hi+lo bf16 operands ("split") if q/k/v or the merger
                  clamped or the shifts run out.  Full speed and max-abs <= 2e-2 against the fp32 reference; never
                  silently clamped.
          "f16"   the same arithmetic without the read: the caller checks `saturation_count()` itself (pipelines, benchmarks).
          "split" hi + lo bf16 operands (~16 mantissa bits, bf16's range, 2x the MMA work); also meets 2e-2.
          "bf16"  activations rounded to bf16 once per GEMM input, as HF's own bf16 path (max-abs ~5e-2 after 32 blocks).
        Default comes from $VR2_PRECISION."""
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.Vr2Error("VisionTower needs a CUDA device: there is no CPU path")
        tc = _lib.TowerConfig()
        for f in ("depth", "hidden_size", "num_heads", "intermediate_size", "out_hidden_size", "patch_size",
                  "temporal_patch_size", "in_channels", "spatial_merge_size", "window_size"):
            setattr(tc, f, int(getattr(cfg, f)))
        full = list(cfg.fullatt_block_indexes)
        if len(full) > 8:
            raise ValueError("at most 8 full-attention layers")
        tc.n_fullatt = len(full)
        for i, v in enumerate(full):
            tc.fullatt_block_indexes[i] = int(v)
        tc.rms_norm_eps = cfg.rms_norm_eps
        tc.rope_theta = cfg.rope_theta
        handle = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(lib().vr2_tower_create(C.byref(tc), index, C.byref(handle)), "vr2_tower_create")
        self._h = handle
        self._plans = collections.OrderedDict()
        self._plan_cache = plan_cache
        self._grid_keys = {}           # id(grid_thw tensor) -> (weakref, version, values): skips the D2H for a tensor seen before
        self._ws = None
        self._ws_split = None          # workspace of the split fallback (twice the size), kept once allocated
        self._emb_cache = collections.OrderedDict()   # embedding reuse, see forward(reuse=...)
        self.reuse_hits = 0
        self.fallbacks = 0             # "auto": forwards repeated in split because the f16 guard tripped
        self.shift_changes = 0         # "auto": times the f16 range shifts were raised (each repeats one forward)
        self.set_precision(precision or os.environ.get("VR2_PRECISION", "auto"))
        self.load_state_dict(state_dict)

    def set_precision(self, precision: str):
        if precision not in _lib.HOST_PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.HOST_PRECISIONS)}, got {precision!r}")
        check(lib().vr2_tower_set_precision(self._h, _lib.HOST_PRECISIONS[precision]), "vr2_tower_set_precision")
        self.precision = precision
        self._ws = None
        self._emb_cache.clear()

    def saturation_count(self, reset: bool = True) -> int:
        """f16 range guard: how many epilogue conversions met a value outside fp16's finite range since the last reset,
        over everything enqueued on the current stream (synchronises it).  0 = nothing was clamped."""
        n = C.c_uint64(0)
        with torch.cuda.device(self.device):
            check(lib().vr2_tower_saturation_count(self._h, C.byref(n), 1 if reset else 0, _stream_ptr(self.device)),
                  "vr2_tower_saturation_count")
        return int(n.value)

    SAT_SITES = ("xb", "qkv", "h", "merger")

    def saturation_counts(self, reset: bool = True) -> dict:
        """The same per conversion site: {"xb": residual-stream operand, "qkv", "h": SwiGLU product, "merger"}."""
        n = (C.c_uint64 * 4)()
        with torch.cuda.device(self.device):
            check(lib().vr2_tower_saturation_counts(self._h, n, 1 if reset else 0, _stream_ptr(self.device)),
                  "vr2_tower_saturation_counts")
        return {k: int(n[i]) for i, k in enumerate(self.SAT_SITES)}

    @property
    def f16_shift(self):
        """(shift_xb, shift_h): the residual-stream operand is stored times 2^-shift_xb and the SwiGLU product times
        2^-shift_h in the f16 / auto precisions (include/vr2.h vr2_tower_set_f16_shift)."""
        v = (C.c_int32 * 2)()
        check(lib().vr2_tower_get_f16_shift(self._h, v), "vr2_tower_get_f16_shift")
        return int(v[0]), int(v[1])

    def set_f16_shift(self, shift_xb: int, shift_h: int):
        check(lib().vr2_tower_set_f16_shift(self._h, int(shift_xb), int(shift_h)), "vr2_tower_set_f16_shift")
        self._emb_cache.clear()

    def saturation_snapshot(self, dst: torch.Tensor):
        """Copies the 4 guard counters into dst[0:4] (int32 tensor, pinned host or device) in stream order, no sync."""
        with torch.cuda.device(self.device):
            check(lib().vr2_tower_saturation_snapshot(self._h, dst.data_ptr(), _stream_ptr(self.device)),
                  "vr2_tower_saturation_snapshot")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().vr2_tower_destroy(h)
            except Exception:
                pass

    def load_state_dict(self, state_dict):
        s = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            for name, t in state_dict.items():
                if name.endswith("inv_freq"):
                    continue  # the library owns its fp32 rotary tables (train_sft.py:46 casts buffers to bf16)
                src = t.detach()
                if src.dtype not in (torch.float32, torch.bfloat16):
                    src = src.float()
                src = src.to(self.device).contiguous()
                shape = (C.c_int64 * max(src.dim(), 1))(*(src.shape if src.dim() else (1,)))
                check(lib().vr2_tower_set_weight(self._h, name.encode(), src.data_ptr(), _dtype_code(src.dtype), shape,
                                                 max(src.dim(), 1), s), f"vr2_tower_set_weight({name})")
                torch.cuda.current_stream(self.device).synchronize()  # src may be a temporary
            check(lib().vr2_tower_finalize(self._h), "vr2_tower_finalize")

    # ------------------------------------------------------------------ plans / workspace
    def plan(self, grid_thw) -> Plan:
        if isinstance(grid_thw, Plan):
            return grid_thw
        if isinstance(grid_thw, torch.Tensor):
            # the values are needed on the host (they size the output): one small D2H, as HF's own `.tolist()`.  The same
            # tensor OBJECT, unmodified, is recognised without it — GRPO passes one `video_grid_thw` through its
            # generate / reference / policy forwards (grpo_trainer.py:783, :826-833, :1094)
            hit = self._grid_keys.get(id(grid_thw))
            if hit is not None and hit[0]() is grid_thw and hit[1] == grid_thw._version:
                key = hit[2]
            else:
                key = tuple(grid_thw.detach().to("cpu", torch.int64).reshape(-1).tolist())
                if len(self._grid_keys) > 64:
                    self._grid_keys.clear()
                self._grid_keys[id(grid_thw)] = (weakref.ref(grid_thw), grid_thw._version, key)
        else:
            key = tuple(int(v) for row in grid_thw for v in row)
        p = self._plans.get(key)
        if p is None:
            p = Plan(self.cfg, [key[i:i + 3] for i in range(0, len(key), 3)], self.device)
            self._plans[key] = p
            while len(self._plans) > self._plan_cache:
                self._plans.popitem(last=False)
        else:
            self._plans.move_to_end(key)
        return p

    def _workspace(self, n_patches: int) -> torch.Tensor:
        need = lib().vr2_tower_workspace_bytes(self._h, n_patches)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, grid_thw, out_dtype=None, out=None, reuse: int = 0,
                check_guard=None) -> torch.Tensor:
        """hidden_states: [N, 1176] fp32/bf16 on this tower's device.  Returns [N/4, out_hidden].

        reuse > 0 keeps the last `reuse` results and returns the cached tensor when the SAME tensor object, unmodified
        (same `_version`), is encoded again with the same grid — a frozen tower encodes one clip several times per
        GRPO step (grpo_trainer.py:667-679 generate, :783 policy and :826-833 reference log-probs).  Identity, not
        content, is the key, so a recycled address can never alias another clip; callers must treat the returned
        tensor as read-only (the reference only reads it: monkey_patch_forward.py:400-406)."""
        plan = self.plan(grid_thw)
        if reuse > 0 and out is None:
            key = (id(hidden_states), plan.grid, out_dtype)
            hit = self._emb_cache.get(key)
            if hit is not None and hit[0]() is hidden_states and hit[1] == hidden_states._version:
                self._emb_cache.move_to_end(key)
                self.reuse_hits += 1
                return hit[2]
            result = self.forward(hidden_states, plan, out_dtype=out_dtype)
            self._emb_cache[key] = (weakref.ref(hidden_states), hidden_states._version, result)
            while len(self._emb_cache) > reuse:
                self._emb_cache.popitem(last=False)
            return result
        x = hidden_states
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.bfloat16)
        x = x.contiguous()
        if x.dim() != 2 or x.shape[1] != self.cfg.patch_dim:
            raise ValueError(f"hidden_states must be [N, {self.cfg.patch_dim}], got {tuple(x.shape)}")
        if x.shape[0] != plan.num_patches:
            raise ValueError(f"hidden_states has {x.shape[0]} rows but grid_thw describes {plan.num_patches} patches")
        odt = out_dtype or x.dtype
        if out is None:
            out = torch.empty((plan.num_tokens, self.cfg.out_hidden_size), dtype=odt, device=self.device)
        self._launch(plan, x, out, None, check_guard)
        return out

    def _enqueue(self, plan, x, out, rows):
        ws = self._workspace(plan.num_patches)
        base = ws.data_ptr()
        aligned = (base + 1023) & ~1023
        with torch.cuda.device(self.device):
            if isinstance(out, (list, tuple)):      # forward_multi: raw destination pointers
                ptrs = (C.c_void_p * len(out))(*[int(v) for v in out])
                check(lib().vr2_tower_forward_multi(self._h, plan._h, x.data_ptr(), _dtype_code(x.dtype), ptrs, len(out),
                                                    aligned, ws.numel() - (aligned - base), _stream_ptr(self.device)),
                      "vr2_tower_forward_multi")
            elif rows is None:
                check(lib().vr2_tower_forward(self._h, plan._h, x.data_ptr(), _dtype_code(x.dtype), out.data_ptr(),
                                              _dtype_code(out.dtype), aligned, ws.numel() - (aligned - base),
                                              _stream_ptr(self.device)), "vr2_tower_forward")
            else:
                check(lib().vr2_tower_forward_scatter(self._h, plan._h, x.data_ptr(), _dtype_code(x.dtype), out.data_ptr(),
                                                      _dtype_code(out.dtype), out.shape[-1], rows.data_ptr(), aligned,
                                                      ws.numel() - (aligned - base), _stream_ptr(self.device)),
                      "vr2_tower_forward_scatter")

    def _launch(self, plan, x, out, rows, check_guard=None):
        """One forward (plain or scattered store).  In the "auto" precision the f16 guard counter is read behind it and
        the call repeated in split when it is non-zero (the same destination rows are simply overwritten)."""
        self._enqueue(plan, x, out, rows)
        if check_guard is None:
            check_guard = self.precision == "auto"
        if not check_guard:
            return
        counts = self.saturation_counts(reset=True)
        while any(counts.values()):
            sx, sh = self.f16_shift
            nx = sx + self.SHIFT_STEP if counts["xb"] else sx
            nh = sh + self.SHIFT_STEP if counts["h"] else sh
            if counts["qkv"] or counts["merger"] or max(nx, nh) > self.SHIFT_MAX:
                self.run_split(plan, x, out, rows)         # nothing a power-of-two shift can absorb
                return
            self.set_f16_shift(nx, nh)
            self.shift_changes += 1
            warnings.warn(f"video_r2_b200: activations left fp16's range at {[k for k, v in counts.items() if v]}; "
                          f"range shifts (xb, h) are now 2^-{nx}, 2^-{nh} and the forward was repeated")
            self._enqueue(plan, x, out, rows)
            counts = self.saturation_counts(reset=True)

    def run_split(self, plan, x, out, rows=None):
        """Repeat a forward in the split precision (fallback of the f16 range guard), then return to the current one."""
        self.fallbacks += 1
        if self.fallbacks == 1:
            warnings.warn("video_r2_b200: activations left fp16's range (saturation guard tripped); this forward was "
                          "repeated with hi+lo bf16 operands.  If it happens on every call, use precision='split'.")
        keep, keep_ws = self.precision, self._ws
        try:
            check(lib().vr2_tower_set_precision(self._h, _lib.PRECISIONS["split"]), "vr2_tower_set_precision")
            self._ws = self._ws_split
            self._enqueue(plan, x, out, rows)
            self._ws_split = self._ws
        finally:
            check(lib().vr2_tower_set_precision(self._h, _lib.HOST_PRECISIONS[keep]), "vr2_tower_set_precision")
            self._ws = keep_ws

    __call__ = forward

    @torch.no_grad()
    def forward_into(self, hidden_states: torch.Tensor, grid_thw, dest: torch.Tensor, dest_rows: torch.Tensor):
        """Tower forward whose merger stores merged token k at dest.view(-1, D)[dest_rows[k]] (dest's dtype): the
        caller's masked_scatter fused into the last GEMM store (include/vr2.h vr2_tower_forward_scatter)."""
        plan = self.plan(grid_thw)
        x = hidden_states
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.bfloat16)
        x = x.contiguous()
        if x.dim() != 2 or x.shape[1] != self.cfg.patch_dim or x.shape[0] != plan.num_patches:
            raise ValueError(f"hidden_states must be [{plan.num_patches}, {self.cfg.patch_dim}], got {tuple(x.shape)}")
        if dest.device != self.device or not dest.is_contiguous() or dest.shape[-1] < self.cfg.out_hidden_size:
            raise ValueError("dest must be a contiguous tensor on the tower's device with rows of >= out_hidden_size")
        rows = dest_rows.to(device=self.device, dtype=torch.int32).contiguous()
        if rows.numel() != plan.num_tokens:
            raise ValueError(f"dest_rows has {rows.numel()} entries for {plan.num_tokens} visual tokens")
        self._launch(plan, x, dest, rows)
        return dest

    @torch.no_grad()
    def forward_multi(self, hidden_states: torch.Tensor, grid_thw, dest_ptrs, check_guard=None):
        """Tower forward whose bf16 result [n_tokens, out_hidden] is stored to every buffer of `dest_ptrs` at once
        (ints: device addresses of row 0 of this call's tokens in each destination, local or peer-mapped memory; at
        most 8) — the all-gather of the sharded multi-GPU mode fused into the merger's last GEMM
        (include/vr2.h vr2_tower_forward_multi).  The caller owns the cross-rank barrier that must follow."""
        plan = self.plan(grid_thw)
        x = hidden_states
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.to(torch.bfloat16)
        x = x.contiguous()
        if x.dim() != 2 or x.shape[1] != self.cfg.patch_dim or x.shape[0] != plan.num_patches:
            raise ValueError(f"hidden_states must be [{plan.num_patches}, {self.cfg.patch_dim}], got {tuple(x.shape)}")
        dest_ptrs = [int(p) for p in dest_ptrs]
        if not 1 <= len(dest_ptrs) <= 8:
            raise ValueError("forward_multi takes 1..8 destinations")
        self._launch(plan, x, dest_ptrs, None, check_guard)

    SHIFT_STEP, SHIFT_MAX = 4, 12

    KERNEL_KINDS = ("cast", "patch_embed", "qkv", "attn_window", "attn_full", "proj", "gate_up", "down",
                    "merger_norm", "fc1", "fc2")

    # ------------------------------------------------------------------ test hooks (stage-by-stage parity)
    def debug_stop(self, n_stages: int):
        """Every following forward returns after n stages (1 = patch embed, then qkv, attention, proj, gate|up, down of
        block 0, ...), leaving the workspace as that stage wrote it; 0 turns it off (include/vr2.h)."""
        check(lib().vr2_tower_set_debug_stop(self._h, int(n_stages)), "vr2_tower_set_debug_stop")

    def debug_buffers(self, n_patches: int) -> dict:
        """Views of the workspace of the LAST forward of n_patches rows (window order): x fp32 [N, hidden], xb / qkv / att
        / h in the activation format of the current precision, ssq fp32 [hidden / 40, N] (partial sums of x^2; how many
        leading rows are live depends on the kernel that wrote them)."""
        off = (C.c_int64 * 7)()
        h_ld = C.c_int64(0)
        check(lib().vr2_tower_workspace_layout(self._h, int(n_patches), off, C.byref(h_ld)), "vr2_tower_workspace_layout")
        ws, n, hd = self._ws, int(n_patches), self.cfg.hidden_size
        act = torch.float16 if self.precision in ("f16", "auto") else torch.bfloat16

        def view(i, dtype, rows, cols):
            nbytes = rows * cols * torch.empty((), dtype=dtype).element_size()
            return ws[off[i]: off[i] + nbytes].view(dtype).reshape(rows, cols)
        return {"x": view(0, torch.float32, n, hd), "xb": view(1, act, n, hd), "qkv": view(2, act, n, 3 * hd),
                "att": view(3, act, n, hd), "h": view(4, act, n, int(h_ld.value)), "ssq": view(5, torch.float32, hd // 40, n)}

    def profile(self, kinds=()):
        """Select kernel kinds to time with CUDA events inside forward (() turns profiling off)."""
        mask = 0
        for k in kinds:
            mask |= 1 << self.KERNEL_KINDS.index(k)
        check(lib().vr2_tower_profile(self._h, mask), "vr2_tower_profile")

    def profile_read(self):
        """-> {kind: (milliseconds, launches)} accumulated since the last read (synchronises)."""
        n = len(self.KERNEL_KINDS)
        ms = (C.c_double * 16)()
        cnt = (C.c_int64 * 16)()
        check(lib().vr2_tower_profile_read(self._h, ms, cnt), "vr2_tower_profile_read")
        return {self.KERNEL_KINDS[i]: (ms[i], cnt[i]) for i in range(n) if cnt[i]}

    def launch_count(self, grid_thw, dtype=torch.bfloat16) -> int:
        return int(lib().vr2_tower_launch_count(self._h, self.plan(grid_thw)._h, _dtype_code(dtype)))
