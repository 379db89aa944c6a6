"""video_r2_b200 — the B200 (sm_100a) implementation of Video-R2's data-parallel hot path:
the Qwen2.5-VL-7B vision tower and the 3-D M-RoPE position indexing, behind the reference's
own plugin mechanism (class-attribute patching of the Hugging Face model classes).

    import video_r2_b200 as vr2
    vr2.install(model)          # model.visual.forward / get_rope_index now run libvr2.so

Everything heavy lives in libvr2.so (video_r2_b200/csrc, C ABI in include/vr2.h)."""
from ._lib import LIB_PATH, Vr2Error, lib
from .rope_index import RopeConfig, get_rope_index
from .embed import embed_visual_tokens, scatter_rows, token_positions
from .logprob import lm_logprob
from .mrope import FusedDecoderQKV, apply_mrope, mrope_table
from .pipeline import HostPipeline, bind_to_gpu_node
from .preprocess import patchify, preprocess_video, resize_frames, smart_nframes, smart_resize, video_target_size
from .tower import Plan, VisionConfig, VisionTower, gather_rows

__all__ = ["LIB_PATH", "Vr2Error", "lib", "RopeConfig", "get_rope_index", "Plan", "VisionConfig", "VisionTower",
           "gather_rows", "HostPipeline", "bind_to_gpu_node", "patchify", "resize_frames", "preprocess_video", "smart_resize", "smart_nframes", "video_target_size", "apply_mrope", "mrope_table", "FusedDecoderQKV", "lm_logprob", "embed_visual_tokens", "scatter_rows", "token_positions", "install", "uninstall", "refresh"]


def install(*args, **kwargs):
    from .patch import install as _install
    return _install(*args, **kwargs)


def uninstall(*args, **kwargs):
    from .patch import uninstall as _uninstall
    return _uninstall(*args, **kwargs)


def refresh(module):
    from .patch import refresh as _refresh
    return _refresh(module)
