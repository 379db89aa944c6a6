"""`get_rope_index` (3-D M-RoPE position ids) on the GPU, same signature and
return values as transformers 4.51.1's
`Qwen2_5_VLForConditionalGeneration.get_rope_index(input_ids, image_grid_thw,
video_grid_thw, second_per_grid_ts, attention_mask)` — the positional call made at
Video-R2 train/src/train/monkey_patch_forward.py:415-421.  No host loop, no
`.tolist()`, no synchronisation: two kernels over device-resident inputs."""
from __future__ import annotations

import ctypes as C
import dataclasses
import os

import torch

from . import _lib
from ._lib import VR2_I32, VR2_I64, VR2_U8, check, lib


@dataclasses.dataclass(frozen=True)
class RopeConfig:
    spatial_merge_size: int = 2
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    tokens_per_second: float = 2
    # True: second_per_grid_t is cast to int64 before the multiply ("normalize type, send to
    # device" lines of transformers >= 4.50; the installed 5.5.0 keeps int()).  False: float32
    # multiply then .long().  See DESIGN.md "open item V1"; override with VR2_SPG_INT_CAST=0/1.
    spg_int_cast: bool = True


def _default_int_cast(cfg: RopeConfig) -> bool:
    env = os.environ.get("VR2_SPG_INT_CAST")
    if env is not None:
        return env not in ("0", "false", "False", "")
    return cfg.spg_int_cast


def get_rope_index(input_ids, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
                   attention_mask=None, *, config: RopeConfig = RopeConfig(), validate: bool = False):
    """Returns (position_ids int64 [3,B,S], mrope_position_deltas int64 [B,1]) on input_ids' device.

    validate=True waits for the kernels and raises ValueError when a row names more images / videos than the grid
    tensors describe (Hugging Face raises IndexError there).  Off by default: the call then never synchronises, and
    such a row is numbered without the missing item (the kernels never read past the grids)."""
    ref = input_ids if input_ids is not None else attention_mask
    if ref is None:
        raise ValueError("get_rope_index needs input_ids or attention_mask")
    if not ref.is_cuda:
        raise _lib.Vr2Error("get_rope_index: tensors must live on a CUDA device (no CPU path)")
    dev = ref.device
    B, S = ref.shape
    ids = None
    if input_ids is not None:
        ids = input_ids.to(torch.int64).contiguous()
    mask, mcode = None, VR2_I64
    if attention_mask is not None:
        mask = attention_mask.to(dev)
        if mask.dtype == torch.int64:
            mcode = VR2_I64
        elif mask.dtype == torch.int32:
            mcode = VR2_I32
        elif mask.dtype in (torch.uint8, torch.bool):
            mcode = VR2_U8
        else:  # float masks: HF compares `== 1`
            mask, mcode = (mask == 1).to(torch.uint8), VR2_U8
        mask = mask.contiguous()

    def grid(g):
        if g is None:
            return None, 0
        g = torch.as_tensor(g, device=dev).to(torch.int64).reshape(-1, 3).contiguous()
        return g, g.shape[0]

    img, n_img = grid(image_grid_thw)
    vid, n_vid = grid(video_grid_thw)
    spg = None
    if second_per_grid_ts is not None:
        # Tensor (GRPO / eval) or list[float] (SFT collator, sft_dataset.py:264-265); torch multiplies in float32
        spg = torch.as_tensor(second_per_grid_ts, dtype=torch.float32).to(dev).reshape(-1).contiguous()
        if vid is not None and spg.numel() < n_vid:
            raise ValueError(f"second_per_grid_ts has {spg.numel()} entries for {n_vid} videos")
    pos = torch.empty((3, B, S), dtype=torch.int64, device=dev)
    deltas = torch.empty((B, 1), dtype=torch.int64, device=dev)
    scratch = torch.empty(lib().vr2_rope_scratch_bytes(B, S), dtype=torch.uint8, device=dev)
    rc = _lib.RopeConfig(config.spatial_merge_size, config.image_token_id, config.video_token_id,
                         config.vision_start_token_id, float(config.tokens_per_second),
                         1 if _default_int_cast(config) else 0)
    p = lambda t: t.data_ptr() if t is not None else None
    with torch.cuda.device(dev):
        check(lib().vr2_rope_index(C.byref(rc), p(ids), p(mask), mcode, B, S, p(img), n_img, p(vid), n_vid, p(spg),
                                   scratch.data_ptr(), pos.data_ptr(), deltas.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream), "vr2_rope_index")
        if validate and ids is not None and (img is not None or vid is not None) and B > 0 and S > 0:
            st = C.c_int32(0)
            check(lib().vr2_rope_index_status(scratch.data_ptr(), B, S, C.byref(st),
                                              torch.cuda.current_stream(dev).cuda_stream), "vr2_rope_index_status")
            if st.value:
                what = " and ".join(n for bit, n in ((1, "image_grid_thw"), (2, "video_grid_thw")) if st.value & bit)
                raise ValueError(f"get_rope_index: input_ids name more vision items than {what} describes")
    return pos, deltas
