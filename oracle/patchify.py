"""ORACLE (test infrastructure, not product code): CPU restatement of the rescale + normalise + patchify of
transformers `Qwen2VLVideoProcessor._preprocess` (video_processing_qwen2_vl.py; rescale_and_normalize and
_fuse_mean_std_and_rescale_factor in image_processing_utils_fast.py), i.e. what produces `pixel_values_videos` for
Video-R2's `self.visual(...)` call (monkey_patch_forward.py:391-392).

    mean' = float32(mean) * (1 / rescale_factor),  std' = float32(std) * (1 / rescale_factor)          (fp32)
    x = (float32(frames) - mean'[c]) / std'[c]                                                         (fp32 sub, then div)
    pad T to a multiple of temporal_patch_size by repeating the last frame
    view [gt, Tp, C, gh/2, 2, P, gw/2, 2, P] -> permute(0, 3, 6, 4, 7, 2, 1, 5, 8) -> [gt*gh*gw, C*Tp*P*P]

Pinned by tests/test_oracle_patchify.py against golden outputs of the HF processor (tests/golden/patchify_*.npz) and
the live processor where importable.  Only tests/, smoke() and bench.py's CPU legs may import this module."""
import numpy as np


def fused_mean_std(image_mean, image_std, rescale_factor=1 / 255):
    inv = np.float32(1.0 / rescale_factor)
    return np.asarray(image_mean, np.float32) * inv, np.asarray(image_std, np.float32) * inv


def patchify(frames, image_mean, image_std, rescale_factor=1 / 255, patch=14, tpatch=2, merge=2):
    """frames: [T, C, H, W] uint8 / float32 numpy -> (float32 [N, C*tpatch*patch*patch], [gt, gh, gw])."""
    x = np.asarray(frames).astype(np.float32)
    mean, std = fused_mean_std(image_mean, image_std, rescale_factor)
    x = (x - mean[None, :, None, None]) / std[None, :, None, None]
    T, C, H, W = x.shape
    pad = -T % tpatch
    if pad:
        x = np.concatenate([x, np.repeat(x[-1:], pad, axis=0)], axis=0)
    gt, gh, gw = x.shape[0] // tpatch, H // patch, W // patch
    x = x.reshape(gt, tpatch, C, gh // merge, merge, patch, gw // merge, merge, patch)
    x = x.transpose(0, 3, 6, 4, 7, 2, 1, 5, 8)
    return np.ascontiguousarray(x.reshape(gt * gh * gw, C * tpatch * patch * patch)), [gt, gh, gw]
