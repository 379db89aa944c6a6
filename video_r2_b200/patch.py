"""Drop-in installation: the same class-attribute patching Video-R2 itself uses to swap
forwards (train/src/train/monkey_patch_forward.py:22-26,
`transformers.models.qwen2_5_vl.modeling_qwen2_5_vl.<Class>.forward = ...`).

    import video_r2_b200 as vr2
    vr2.install()            # before or after from_pretrained; callers stay unchanged

After install():
  * `Qwen2_5_VisionTransformerPretrainedModel.forward(hidden_states, grid_thw)` runs libvr2.so.
    The first call on a module instance snapshots its weights (`state_dict()`, all-gathered first when
    DeepSpeed ZeRO-3 has partitioned them: `full_parameters`) into library-owned, repacked storage — valid because every shipped Video-R2 recipe freezes the tower
    (train_sft.sh:42-44, train_grpo.sh:47-49); call `vr2.refresh(module)` after changing weights.
    Return type follows the installed transformers: a bare Tensor on 4.x (what
    monkey_patch_forward.py:392 expects), `BaseModelOutputWithPooling(pooler_output=...)` on >= 5.
  * `get_rope_index(input_ids, image_grid_thw, video_grid_thw, second_per_grid_ts, attention_mask)`
    (4.51.1 positional order, monkey_patch_forward.py:415-421) runs on the GPU.  It is attached to
    whichever class owns it in the installed version (`...ForConditionalGeneration` on 4.51,
    `Qwen2_5_VLModel` on >= 4.52); the 5.x variant with `mm_token_type_ids` is accepted and that
    argument ignored — the semantics implemented are 4.51.1's, the version the reference pins.
"""
from __future__ import annotations

import contextlib

import torch

from .rope_index import RopeConfig, get_rope_index
from .tower import VisionConfig, VisionTower

_ORIGINALS = {}
_REUSE = 0   # install(reuse_embeddings=n): keep the last n tower outputs per module (see VisionTower.forward)
_PRECISION = None   # install(precision=...): operand precision of the towers built from now on (None: $VR2_PRECISION / auto)


def _is_zero3_partitioned(p) -> bool:
    """DeepSpeed ZeRO-3 (`zero.Init`, what `from_pretrained` runs under with zero3.json) replaces every parameter's
    storage with an empty placeholder and hangs the rank's shard off `ds_tensor`; `ds_id` marks such a parameter."""
    return hasattr(p, "ds_id") or hasattr(p, "ds_tensor")


@contextlib.contextmanager
def full_parameters(module):
    """Context in which `module.state_dict()` holds the FULL weights.  A no-op for ordinary modules.  Under ZeRO-3
    (both shipped Video-R2 recipes: train/scripts/train_sft.sh:29, train_grpo.sh:42 -> zero3.json:16-27) the
    parameters of every sub-module are partitioned and DeepSpeed's pre-forward hook gathers only the parameters of
    the module whose `forward` runs -- this patch replaces the tower's forward wholesale, so the blocks' hooks never
    fire; all of the tower's parameters are therefore all-gathered here once (a collective: every rank reaches the
    first `self.visual(...)` call, which is what the reference's dummy-image path guarantees,
    monkey_patch_forward.py:358-369), copied into the library's own storage, and released again."""
    params = [p for p in module.parameters() if _is_zero3_partitioned(p)]
    if not params:
        yield
        return
    try:
        import deepspeed
    except ImportError as e:  # partitioned parameters without deepspeed cannot be gathered: say so instead of
        raise RuntimeError(   # handing empty placeholders to the library
            "video_r2_b200: the vision tower's parameters are ZeRO-3 partitioned (ds_id present) but `deepspeed` "
            "is not importable, so they cannot be gathered") from e
    with deepspeed.zero.GatheredParameters(params, modifier_rank=None):
        yield


def _tower_of(module) -> VisionTower:
    tw = module.__dict__.get("_vr2_tower")
    dev = next(module.parameters()).device
    if tw is None or tw.device != dev:
        cfg = VisionConfig.from_hf(module.config)
        with full_parameters(module):
            # VisionTower copies (and repacks) every tensor into library-owned storage before returning
            tw = VisionTower(cfg, module.state_dict(), device=dev, precision=_PRECISION)
        module.__dict__["_vr2_tower"] = tw
    return tw


def refresh(module):
    """Re-read the module's weights (after loading a checkpoint or an optimizer step on the tower)."""
    module.__dict__.pop("_vr2_tower", None)


def _transformers_major() -> int:
    import transformers
    return int(transformers.__version__.split(".")[0])


def _visual_forward(self, hidden_states: torch.Tensor, grid_thw: torch.Tensor, **kwargs):
    out = _tower_of(self).forward(hidden_states, grid_thw, reuse=_REUSE)
    if _transformers_major() >= 5:
        from transformers.modeling_outputs import BaseModelOutputWithPooling
        return BaseModelOutputWithPooling(last_hidden_state=None, pooler_output=out)
    return out


def _rope_cfg(self) -> RopeConfig:
    cfg = self.config
    vis = getattr(cfg, "vision_config", None)
    return RopeConfig(
        spatial_merge_size=getattr(vis, "spatial_merge_size", 2),
        image_token_id=getattr(cfg, "image_token_id", 151655),
        video_token_id=getattr(cfg, "video_token_id", 151656),
        vision_start_token_id=getattr(cfg, "vision_start_token_id", 151652),
        tokens_per_second=getattr(vis, "tokens_per_second", 2),
    )


def _rope_4x(self, input_ids=None, image_grid_thw=None, video_grid_thw=None, second_per_grid_ts=None,
             attention_mask=None):
    return get_rope_index(input_ids, image_grid_thw, video_grid_thw, second_per_grid_ts, attention_mask,
                          config=_rope_cfg(self))


def _rope_5x(self, input_ids=None, mm_token_type_ids=None, image_grid_thw=None, video_grid_thw=None,
             second_per_grid_ts=None, attention_mask=None, **kwargs):
    return get_rope_index(input_ids, image_grid_thw, video_grid_thw, second_per_grid_ts, attention_mask,
                          config=_rope_cfg(self))


def install(visual: bool = True, rope_index: bool = True, reuse_embeddings: int = 0, precision: str = None):
    """Patch the Hugging Face classes in place.  Idempotent.  reuse_embeddings=n keeps the last n tower outputs and
    returns them when the same (unmodified) pixel tensor object is encoded again — GRPO encodes each clip three times
    per step with a frozen tower; off by default because the returned tensor is then shared between calls.
    precision: "auto" | "f16" | "split" | "bf16" for the towers built after this call (see VisionTower; default
    $VR2_PRECISION, else "auto": fp16 activations with the saturation guard and an automatic split re-run, the mode
    that meets the north-star tolerance max-abs 2e-2 / cosine 0.9995 against the fp32 reference)."""
    from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as M
    from . import _lib, lib
    global _REUSE, _PRECISION
    if precision is not None and precision not in _lib.HOST_PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(_lib.HOST_PRECISIONS)}, got {precision!r}")
    _REUSE = int(reuse_embeddings)
    _PRECISION = precision
    lib()  # fail now, loudly, if libvr2.so is missing
    if visual and "visual" not in _ORIGINALS:
        cls = M.Qwen2_5_VisionTransformerPretrainedModel
        _ORIGINALS["visual"] = (cls, cls.forward)
        cls.forward = _visual_forward
    if rope_index and "rope" not in _ORIGINALS:
        fn = _rope_5x if _transformers_major() >= 5 else _rope_4x
        saved = []
        for name in ("Qwen2_5_VLModel", "Qwen2_5_VLForConditionalGeneration"):
            cls = getattr(M, name, None)
            if cls is not None and "get_rope_index" in cls.__dict__:
                saved.append((cls, cls.__dict__["get_rope_index"]))
                cls.get_rope_index = fn
        _ORIGINALS["rope"] = saved


def uninstall():
    if "visual" in _ORIGINALS:
        cls, fwd = _ORIGINALS.pop("visual")
        cls.forward = fwd
    for cls, fn in _ORIGINALS.pop("rope", []):
        cls.get_rope_index = fn
