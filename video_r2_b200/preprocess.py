"""Frames -> the tower's input rows on the GPU: the rescale + normalise + patchify of Hugging Face's
`Qwen2VLVideoProcessor._preprocess` (transformers video_processing_qwen2_vl.py) and the caller's cast to the tower
dtype (Video-R2 monkey_patch_forward.py:391), in one kernel (`vr2_patchify`).  The resize stays where Video-R2 does it
(train/src/dataset/vision_process.py: torchvision bicubic on the CPU); what changes is that the host ships uint8 frames
(1 byte per pixel value) instead of fp32 patch rows (4 bytes)."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import VR2_BF16, VR2_F32, VR2_U8, check, lib
from .tower import _stream_ptr

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def fused_mean_std(image_mean, image_std, rescale_factor: float = 1 / 255, do_rescale: bool = True):
    """HF's _fuse_mean_std_and_rescale_factor: fp32 tensors times the Python scalar 1/rescale_factor."""
    mean = torch.tensor(image_mean, dtype=torch.float32)
    std = torch.tensor(image_std, dtype=torch.float32)
    if do_rescale:
        mean = mean * (1.0 / rescale_factor)
        std = std * (1.0 / rescale_factor)
    return mean, std


@torch.no_grad()
def patchify(frames: torch.Tensor, image_mean=OPENAI_CLIP_MEAN, image_std=OPENAI_CLIP_STD, rescale_factor: float = 1 / 255,
             patch_size: int = 14, temporal_patch_size: int = 2, merge_size: int = 2, out_dtype=torch.bfloat16,
             out: torch.Tensor = None):
    """frames: [T, C, H, W] uint8 or float32 CUDA tensor, already resized (H, W multiples of patch*merge).
    -> (pixel_values [N, C*Tp*P*P] of out_dtype, grid_thw [t, h, w]) exactly as the HF processor + `.type(dtype)`."""
    if not frames.is_cuda or frames.dim() != 4:
        raise ValueError("patchify: frames must be a [T, C, H, W] CUDA tensor (no CPU path)")
    if frames.dtype not in (torch.uint8, torch.float32):
        raise ValueError(f"patchify: frames must be uint8 or float32, got {frames.dtype}")
    frames = frames.contiguous()
    T, Cc, H, W = frames.shape
    mean, std = fused_mean_std(image_mean, image_std, rescale_factor)
    gt = -(-T // temporal_patch_size)
    if H % (patch_size * merge_size) or W % (patch_size * merge_size):
        raise ValueError(f"patchify: frame size {H}x{W} is not a multiple of patch*merge = {patch_size * merge_size}")
    gh, gw = H // patch_size, W // patch_size
    rows, cols = gt * gh * gw, Cc * temporal_patch_size * patch_size * patch_size
    if out is None:
        out = torch.empty((rows, cols), dtype=out_dtype, device=frames.device)
    elif tuple(out.shape) != (rows, cols) or not out.is_contiguous():
        raise ValueError(f"patchify: out must be a contiguous [{rows}, {cols}] tensor")
    m = (C.c_float * Cc)(*mean.tolist())
    s = (C.c_float * Cc)(*std.tolist())
    with torch.cuda.device(frames.device):
        check(lib().vr2_patchify(frames.data_ptr(), VR2_U8 if frames.dtype == torch.uint8 else VR2_F32, T, Cc, H, W,
                                 patch_size, temporal_patch_size, merge_size, m, s, out.data_ptr(),
                                 VR2_BF16 if out.dtype == torch.bfloat16 else VR2_F32, _stream_ptr(frames.device)),
              "vr2_patchify")
    return out, [gt, gh, gw]
