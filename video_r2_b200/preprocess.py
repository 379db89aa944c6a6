"""Frames -> the tower's input rows on the GPU: the rescale + normalise + patchify of Hugging Face's
`Qwen2VLVideoProcessor._preprocess` (transformers video_processing_qwen2_vl.py) and the caller's cast to the tower
dtype (Video-R2 monkey_patch_forward.py:391), in one kernel (`vr2_patchify`); and, before it, the anti-aliased bicubic
resize Video-R2's `fetch_video` runs on the host cores (train/qwen_vl_utils/vision_process.py:771-810) as
`vr2_resize_frames`, with the integer helpers that choose the target size and frame count (`smart_resize` :336-362,
`smart_nframes` :425-462) restated for callers that no longer import qwen_vl_utils.  The host ships the decoder's
uint8 frames at source resolution (1 byte per pixel value) instead of fp32 patch rows (4 bytes); both kernels are
byte-identical to the CPU path they replace."""
from __future__ import annotations

import ctypes as C
import math

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


# ---------------------------------------------------------------------------------------------
# target size / frame count (host integers) and the resize itself
# ---------------------------------------------------------------------------------------------
IMAGE_FACTOR = 28
VIDEO_MIN_PIXELS = 128 * 28 * 28
VIDEO_MAX_PIXELS = 768 * 28 * 28
VIDEO_TOTAL_PIXELS = int(128000 * 28 * 28 * 0.9)
FRAME_FACTOR = 2
MAX_RATIO = 200


def smart_resize(height: int, width: int, factor: int = IMAGE_FACTOR, min_pixels: int = 4 * 28 * 28,
                 max_pixels: int = 16384 * 28 * 28):
    """Target (h, w): multiples of `factor`, area within [min_pixels, max_pixels], aspect ratio kept as closely as the
    grid allows.  Same results as vision_process.py:336-362 (round half to even; floor when shrinking, ceil when growing)."""
    if max(height, width) / min(height, width) > MAX_RATIO:
        raise ValueError(f"absolute aspect ratio must be smaller than {MAX_RATIO}, got {max(height, width) / min(height, width)}")
    h = max(factor, round(height / factor) * factor)
    w = max(factor, round(width / factor) * factor)
    if h * w > max_pixels:
        shrink = math.sqrt((height * width) / max_pixels)
        h = math.floor(height / shrink / factor) * factor
        w = math.floor(width / shrink / factor) * factor
    elif h * w < min_pixels:
        grow = math.sqrt(min_pixels / (height * width))
        h = math.ceil(height * grow / factor) * factor
        w = math.ceil(width * grow / factor) * factor
    return h, w


def smart_nframes(total_frames: int, video_fps: float, fps: float = 2.0, min_frames: int = 4, max_frames: int = 32,
                  nframes: int = None) -> int:
    """How many frames to sample (an even count).  vision_process.py:425-462 with the `ele` dict spelled out as arguments
    (`max_frames` defaults to the reference's FPS_MAX_FRAMES = 32; eval sets 128)."""
    if nframes is not None:
        n = round(nframes / FRAME_FACTOR) * FRAME_FACTOR
    else:
        lo = math.ceil(min_frames / FRAME_FACTOR) * FRAME_FACTOR
        hi = math.floor(min(max_frames, total_frames) / FRAME_FACTOR) * FRAME_FACTOR
        n = min(min(max(total_frames / video_fps * fps, lo), hi), total_frames)
        n = math.floor(n / FRAME_FACTOR) * FRAME_FACTOR
    if not (FRAME_FACTOR <= n <= total_frames):
        raise ValueError(f"nframes should in interval [{FRAME_FACTOR}, {total_frames}], but got {n}.")
    return int(n)


def video_target_size(nframes: int, height: int, width: int, min_pixels: int = VIDEO_MIN_PIXELS,
                      total_pixels: int = VIDEO_TOTAL_PIXELS, max_pixels: int = None, factor: int = IMAGE_FACTOR):
    """fetch_video's choice of (h', w') for a clip of `nframes` frames (vision_process.py:780-801)."""
    cap = max(min(VIDEO_MAX_PIXELS, total_pixels / nframes * FRAME_FACTOR), int(min_pixels * 1.05))
    if max_pixels is not None:
        cap = min(max_pixels, cap)
    return smart_resize(height, width, factor=factor, min_pixels=min_pixels, max_pixels=cap)


_resize_ws = {}


@torch.no_grad()
def resize_frames(frames: torch.Tensor, size, out_dtype=torch.uint8, out: torch.Tensor = None) -> torch.Tensor:
    """frames: uint8 CUDA tensor [T, C, H, W] with any strides (a permuted [T, H, W, C] decoder buffer is fine);
    size: (h', w').  -> contiguous [T, C, h', w'] uint8, or float32 holding the same integers (what the reference's
    `.float()` returns).  Anti-aliased bicubic, byte-identical to torchvision's CPU tensor path."""
    if not frames.is_cuda or frames.dim() != 4 or frames.dtype != torch.uint8:
        raise ValueError("resize_frames: frames must be a uint8 [T, C, H, W] CUDA tensor (no CPU path)")
    if out_dtype not in (torch.uint8, torch.float32):
        raise ValueError("resize_frames: out_dtype must be uint8 or float32")
    T, Cc, H, W = frames.shape
    oh, ow = int(size[0]), int(size[1])
    if any(st < 0 for st in frames.stride()):
        frames = frames.contiguous()
    if out is None:
        out = torch.empty((T, Cc, oh, ow), dtype=out_dtype, device=frames.device)
    elif tuple(out.shape) != (T, Cc, oh, ow) or not out.is_contiguous() or out.dtype != out_dtype:
        raise ValueError(f"resize_frames: out must be a contiguous {out_dtype} [{T}, {Cc}, {oh}, {ow}] tensor")
    nbytes = C.c_size_t()
    check(lib().vr2_resize_workspace_bytes(H, W, oh, ow, C.byref(nbytes)), "vr2_resize_workspace_bytes")
    key = (frames.device, torch.cuda.current_stream(frames.device).cuda_stream)
    ws = _resize_ws.get(key)                     # per (device, stream): stream order protects the tables between calls
    if ws is None or ws.numel() < nbytes.value:
        ws = torch.empty(max(nbytes.value, 1 << 16), dtype=torch.uint8, device=frames.device)
        _resize_ws[key] = ws
    st, sc, sh, sw = frames.stride()
    planes_per_call = max(1, 65535 // Cc)
    with torch.cuda.device(frames.device):
        for t0 in range(0, T, planes_per_call):
            n = min(planes_per_call, T - t0)
            check(lib().vr2_resize_frames(frames[t0:].data_ptr(), n, Cc, H, W, st, sc, sh, sw, oh, ow, out[t0:].data_ptr(),
                                          VR2_U8 if out_dtype == torch.uint8 else VR2_F32, ws.data_ptr(),
                                          _stream_ptr(frames.device)), "vr2_resize_frames")
    return out


@torch.no_grad()
def preprocess_video(frames: torch.Tensor, size=None, min_pixels: int = VIDEO_MIN_PIXELS, total_pixels: int = VIDEO_TOTAL_PIXELS,
                     max_pixels: int = None, out_dtype=torch.bfloat16, **patchify_kw):
    """Decoder frames (uint8 CUDA [T, C, H, W], source resolution) -> (pixel_values [N, 1176], grid_thw [t, h, w]):
    fetch_video's resize followed by the HF processor's rescale / normalise / patchify, two kernels."""
    T, _, H, W = frames.shape
    if size is None:
        size = video_target_size(T, H, W, min_pixels=min_pixels, total_pixels=total_pixels, max_pixels=max_pixels)
    resized = frames if (H, W) == tuple(size) and frames.is_contiguous() else resize_frames(frames, size)
    return patchify(resized, out_dtype=out_dtype, **patchify_kw)
