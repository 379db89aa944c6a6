"""TEST INFRASTRUCTURE — CPU restatement of the frame resize Video-R2 does before the HF processor
(ref: train/qwen_vl_utils/vision_process.py:771-810 `fetch_video`, identical copies under demo/ and eval/):

    video = transforms.functional.resize(video_u8[T,3,H,W], [h', w'], BICUBIC, antialias=True).float()

and of the integer helpers that pick h', w' and the frame count (`smart_resize` :336-362, `smart_nframes` :425-462).

torchvision's tensor path (torchvision/transforms/_functional_tensor.py `resize`) casts uint8 to float32, calls
`torch.nn.functional.interpolate(mode="bicubic", antialias=True, align_corners=False)`, clamps to [0, 255], rounds
half-to-even and casts back.  The arithmetic lives in ATen (not vendored in /root/reference; torch 2.7 pinned by
ref:environment.yml, torch 2.11.0 in this image): a separable filter, width first then height, Keys cubic with
a = -0.5, support 2*max(scale,1), weights normalised per output index, everything in float32.

Pinned: against `torch.nn.functional.interpolate` of this image, bit for bit on the float32 result (before the
rounding), on random sizes and scales up to 72 taps (tests/test_oracle_resize.py + tests/golden/resize_*.npz, generated
by tests/golden/make_golden.py).  Two things had to be read off the library's behaviour rather than its source:
  * the cubic polynomial and the products are contracted to FMAs by the compiler (`cubic`, below);
  * the tap loop `acc = x0*w0; for j in 1..n-1: acc += xj*wj` is vectorised as an in-order reduction: the first
    4*floor((n-1)/4) taps after the peeled one are multiply-then-add, the remaining (n-1)%4 are fused multiply-adds.
They only matter on exact .5 ties of the final rounding, but the GPU kernel follows them so the bytes are identical.

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module."""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
f64 = np.float64

IMAGE_FACTOR = 28
VIDEO_MIN_PIXELS = 128 * 28 * 28
VIDEO_MAX_PIXELS = 768 * 28 * 28
VIDEO_TOTAL_PIXELS_DEFAULT = None      # the reference reads it from an env-overridable constant; callers pass it
FRAME_FACTOR = 2
MAX_RATIO = 200


# ------------------------------------------------------------------ integer helpers (vision_process.py:320-362, :425-462)
def smart_resize(height, width, factor=IMAGE_FACTOR, min_pixels=4 * 28 * 28, max_pixels=16384 * 28 * 28):
    if max(height, width) / min(height, width) > MAX_RATIO:
        raise ValueError("aspect ratio above %d" % MAX_RATIO)
    hb = max(factor, round(height / factor) * factor)        # Python round(): half to even
    wb = max(factor, round(width / factor) * factor)
    if hb * wb > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        hb = math.floor(height / beta / factor) * factor
        wb = math.floor(width / beta / factor) * factor
    elif hb * wb < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        hb = math.ceil(height * beta / factor) * factor
        wb = math.ceil(width * beta / factor) * factor
    return hb, wb


def smart_nframes(total_frames, video_fps, fps=2.0, min_frames=4, max_frames=32, nframes=None):
    if nframes is not None:
        n = round(nframes / FRAME_FACTOR) * FRAME_FACTOR
    else:
        lo = math.ceil(min_frames / FRAME_FACTOR) * FRAME_FACTOR
        hi = math.floor(min(max_frames, total_frames) / FRAME_FACTOR) * FRAME_FACTOR
        n = total_frames / video_fps * fps
        n = min(min(max(n, lo), hi), total_frames)
        n = math.floor(n / FRAME_FACTOR) * FRAME_FACTOR
    if not (FRAME_FACTOR <= n <= total_frames):
        raise ValueError(f"nframes should be in [{FRAME_FACTOR}, {total_frames}], got {n}")
    return int(n)


def video_max_pixels(nframes, min_pixels=VIDEO_MIN_PIXELS, total_pixels=None, max_pixels=None):
    """fetch_video :781-787: the per-frame pixel cap."""
    if total_pixels is None:
        total_pixels = int(float(128000 * 28 * 28 * 0.9))
    cap = max(min(VIDEO_MAX_PIXELS, total_pixels / nframes * FRAME_FACTOR), int(min_pixels * 1.05))
    return min(max_pixels, cap) if max_pixels is not None else cap


# ------------------------------------------------------------------ float32 arithmetic, spelled out
def _fma(a, b, c):
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)   # 24x24-bit product is exact in f64


def _mul(a, b):
    return (np.asarray(a, f32) * np.asarray(b, f32)).astype(f32)


def _add(a, b):
    return (np.asarray(a, f32) + np.asarray(b, f32)).astype(f32)


def cubic(x):
    """Keys kernel, a = -0.5, Horner form with every multiply-add fused."""
    x = np.abs(np.asarray(x, f32))
    near = _fma(_mul(_fma(f32(1.5), x, f32(-2.5)), x), x, f32(1))                 # ((a+2)x - (a+3)) x x + 1
    far = _fma(_fma(_fma(f32(-0.5), x, f32(2.5)), x, f32(-4)), x, f32(2))          # ((a x - 5a) x + 8a) x - 4a
    return np.where(x < 1, near, np.where(x < 2, far, f32(0))).astype(f32)


def axis_weights(in_size: int, out_size: int):
    """-> list over output index of (first input index, float32 weights[n_taps])."""
    scale = f32(f32(in_size) / f32(out_size))
    support = f32(2) * scale if scale >= 1 else f32(2)
    invscale = f32(1.0 / f64(scale)) if scale >= 1 else f32(1)
    max_taps = int(math.ceil(support)) * 2 + 1
    table = []
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        lo = max(int(f64(f32(center - support)) + 0.5), 0)                         # int64 cast: toward zero
        n = min(int(f64(f32(center + support)) + 0.5), in_size) - lo
        n = min(max(n, 0), max_taps)
        j = np.arange(n)
        arg = (f64(1) * ((j + lo).astype(f32) - center).astype(f64) + 0.5) * f64(invscale)
        w = cubic(arg.astype(f32))
        tot = f32(0)
        for v in w:
            tot = f32(tot + v)
        if tot != 0:
            w = (w / tot).astype(f32)
        table.append((lo, w))
    return table


def resample_last_axis(x: np.ndarray, out_size: int) -> np.ndarray:
    """x [..., in] float32 -> [..., out] float32, with the library's tap order (module docstring)."""
    table = axis_weights(x.shape[-1], out_size)
    out = np.zeros(x.shape[:-1] + (out_size,), f32)
    for i, (lo, w) in enumerate(table):
        n = len(w)
        if n == 0:
            continue
        acc = _mul(x[..., lo], w[0])
        plain = (n - 1) // 4 * 4
        for j in range(1, n):
            acc = _add(acc, _mul(x[..., lo + j], w[j])) if j <= plain else _fma(x[..., lo + j], w[j], acc)
        out[..., i] = acc
    return out


def resize_bicubic_aa_f32(frames: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[..., H, W] (any real dtype) -> float32 [..., out_h, out_w]: F.interpolate(bicubic, antialias) before rounding."""
    x = np.asarray(frames).astype(f32)
    if out_w != x.shape[-1]:
        x = resample_last_axis(x, out_w)
    if out_h != x.shape[-2]:
        x = np.swapaxes(resample_last_axis(np.swapaxes(x, -1, -2), out_h), -1, -2)
    return np.ascontiguousarray(x)


def resize_frames(frames_u8: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [T, C, H, W] -> uint8 [T, C, out_h, out_w]: clamp, round half to even, cast (torchvision's uint8 path)."""
    y = resize_bicubic_aa_f32(frames_u8, out_h, out_w)
    return np.rint(np.clip(y, 0, 255)).astype(np.uint8)
