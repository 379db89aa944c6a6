"""Generates the golden fixtures in this directory by running the REFERENCE arithmetic —
the Hugging Face `Qwen2_5_VisionTransformerPretrainedModel` that Video-R2 calls
(train/src/train/monkey_patch_forward.py:392) — in the build container:

    python tests/golden/make_golden.py

transformers 5.5.0 is what is installed here (Video-R2 pins 4.51.1; SURVEY §8c shows
the vision-tower math is unchanged between them, only the return type is).  The
fixtures pin oracle/indices.py and oracle/tower.py; the GPU tests then compare the
CUDA path with the oracle and with these files.

  indices_<name>.npz  window_index, raw + deduped cu_window_seqlens, rot_pos_emb output of
                      HF get_window_index / rot_pos_emb
  tower_<name>.npz    fp32 `.pooler_output` of the HF module (attn_implementation="sdpa") on
                      seeded weights (oracle.tower.make_weights) and seeded pixels; inputs are
                      regenerated from the seeds, their checksums are stored
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import tower as otower  # noqa: E402

INDEX_GRIDS = {
    "docstring_4x12x20": [[4, 12, 20]],
    "cfg1_8x32x32": [[8, 32, 32]],
    "zero3_dummy_1x28x28": [[1, 28, 28]],
    "tiny_1x2x2": [[1, 2, 2]],
    "odd_merged_3x26x46": [[3, 26, 46]],
    "multi_clip": [[2, 10, 14], [1, 4, 6], [3, 26, 46], [1, 8, 8]],
    "cfg4_slice_2x24x42": [[2, 24, 42]],
}

TOWER_CASES = {
    # name: (depth, fullatt, grid, weight seed, pixel seed)
    "depth32_ragged": (32, (7, 15, 23, 31), [[2, 8, 12], [1, 4, 6]], 11, 12),
    "depth4_odd": (4, (1, 3), [[1, 10, 14], [2, 6, 4]], 21, 22),
}


def hf_model(depth, full):
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel as V

    cfg = Qwen2_5_VLVisionConfig(depth=depth, hidden_size=1280, intermediate_size=3420, num_heads=16,
                                 out_hidden_size=3584, patch_size=14, spatial_merge_size=2, temporal_patch_size=2,
                                 window_size=112, fullatt_block_indexes=list(full), tokens_per_second=2,
                                 hidden_act="silu")
    cfg._attn_implementation = "sdpa"
    return V._from_config(cfg).eval().float()


def seeded_pixels(n, seed):
    x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(seed))
    return x.to(torch.bfloat16).to(torch.float32)  # bf16-exact, like normalised frames cast by the caller


def checksum(t: torch.Tensor) -> np.ndarray:
    t = t.double().reshape(-1)
    return np.array([t.sum().item(), t.abs().sum().item(), t[:: max(1, t.numel() // 97)].sum().item()])


def main():
    m = hf_model(1, [0])
    for name, grid in INDEX_GRIDS.items():
        g = torch.tensor(grid)
        wi, cu = m.get_window_index(g)
        dedup = torch.unique_consecutive(torch.tensor(cu, dtype=torch.int32))
        rot = m.rot_pos_emb(g)
        np.savez_compressed(os.path.join(HERE, f"indices_{name}.npz"), grid=np.asarray(grid, dtype=np.int64),
                            window_index=wi.numpy().astype(np.int64), cu_raw=np.asarray(cu, dtype=np.int64),
                            cu_dedup=dedup.numpy().astype(np.int32), rot_pos_emb=rot.numpy().astype(np.float32))
        print("indices", name, wi.shape, len(cu), dedup.shape)
    for name, (depth, full, grid, wseed, pseed) in TOWER_CASES.items():
        cfg = otower.TowerConfig(depth=depth, fullatt_block_indexes=tuple(full))
        w = otower.make_weights(cfg, seed=wseed)
        m = hf_model(depth, full)
        m.load_state_dict(w, strict=True)
        n = sum(t * h * ww for t, h, ww in grid)
        x = seeded_pixels(n, pseed)
        with torch.no_grad():
            res = m(x, grid_thw=torch.tensor(grid))
        out = res.pooler_output if hasattr(res, "pooler_output") else res
        wsum = np.stack([checksum(w[k]) for k in sorted(w)][:8])
        np.savez_compressed(os.path.join(HERE, f"tower_{name}.npz"), grid=np.asarray(grid, dtype=np.int64),
                            depth=depth, fullatt=np.asarray(full), weight_seed=wseed, pixel_seed=pseed,
                            out=out.numpy().astype(np.float32), pixel_checksum=checksum(x), weight_checksum=wsum)
        print("tower", name, out.shape, float(out.abs().max()))


PATCHIFY_CASES = {"even_4x56x84": ((4, 56, 84), 0), "odd_5x28x56": ((5, 28, 56), 1), "single_1x28x28": ((1, 28, 28), 2)}


def patchify_goldens():
    """uint8 frames -> pixel_values_videos through the HF video processor itself (resize off)."""
    from transformers.models.qwen2_vl.video_processing_qwen2_vl import Qwen2VLVideoProcessor
    proc = Qwen2VLVideoProcessor(do_resize=False)
    for name, ((T, H, W), seed) in PATCHIFY_CASES.items():
        v = torch.randint(0, 256, (T, 3, H, W), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
        out = proc(videos=[v], return_tensors="pt", do_resize=False, do_sample_frames=False)
        np.savez_compressed(os.path.join(HERE, f"patchify_{name}.npz"), frames=v.numpy(),
                            out=out["pixel_values_videos"].numpy(), grid=out["video_grid_thw"].numpy()[0],
                            mean=np.asarray(proc.image_mean), std=np.asarray(proc.image_std), rescale=proc.rescale_factor)
        print("patchify", name, tuple(out["pixel_values_videos"].shape))


def mrope_goldens():
    """q, k, position ids -> rotated q, k through HF's own rotary code (fp32 and bf16)."""
    from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as M
    g = torch.Generator().manual_seed(5)
    B, S, Hq, Hk, D = 2, 19, 3, 2, 128
    pos = torch.randint(0, 5000, (3, B, S), generator=g)
    pos[:, 1, :4] = 1                                   # padding positions as get_rope_index writes them
    inv = 1.0 / (1000000.0 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    out = {"pos": pos.numpy(), "inv_freq": inv.numpy(), "section": np.asarray([16, 24, 24])}
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        q = torch.randn(B, Hq, S, D, generator=g).to(dt)
        k = torch.randn(B, Hk, S, D, generator=g).to(dt)
        inv_e = inv[None, None, :, None].float().expand(3, B, -1, 1)      # Qwen2_5_VLRotaryEmbedding.forward
        freqs = (inv_e @ pos[:, :, None, :].float()).transpose(2, 3)
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos().to(dt), emb.sin().to(dt)
        rq, rk = M.apply_multimodal_rotary_pos_emb(q, k, cos, sin, [16, 24, 24])
        for key, t in (("q", q), ("k", k), ("rq", rq), ("rk", rk)):
            out[f"{key}_{name}"] = t.float().numpy()      # bf16 values are exactly representable in fp32
    np.savez_compressed(os.path.join(HERE, "mrope_small.npz"), **out)
    print("mrope", out["rq_bf16"].shape)


RESIZE_CASES = {  # name: ((T, H, W), (h', w'), channels_last source, seed)
    "down_90x160_to_56x84": ((2, 90, 160), (56, 84), False, 0),
    "up_50x70_to_84x112": ((1, 50, 70), (84, 112), False, 1),
    "hwc_72x128_to_28x56": ((3, 72, 128), (28, 56), True, 2),        # decord layout: [T,H,W,C] permuted to [T,C,H,W]
    "steep_200x37_to_28x28": ((1, 200, 37), (28, 28), False, 3),     # 29 taps down, 5 taps sideways
    "width_only_64x100_to_64x56": ((1, 64, 100), (64, 56), False, 4),
}


def resize_goldens():
    """uint8 frames through the reference's own call (vision_process.py:802-807: torchvision tensor resize, bicubic,
    antialias) plus the float32 result of F.interpolate before the rounding; and tables of the reference module's
    smart_resize / smart_nframes (imported from /root/reference) over a sweep of arguments."""
    import torch.nn.functional as F
    from torchvision import transforms
    from torchvision.transforms import InterpolationMode
    for name, ((T, H, W), (oh, ow), hwc, seed) in RESIZE_CASES.items():
        g = torch.Generator().manual_seed(seed)
        if hwc:
            v = torch.randint(0, 256, (T, H, W, 3), dtype=torch.uint8, generator=g).permute(0, 3, 1, 2)
        else:
            v = torch.randint(0, 256, (T, 3, H, W), dtype=torch.uint8, generator=g)
        v[0, :, : H // 2] = (v[0, :, : H // 2].float() * 0.1 + 235).to(torch.uint8)   # a bright region: overshoot -> clamp at 255
        out = transforms.functional.resize(v, [oh, ow], interpolation=InterpolationMode.BICUBIC, antialias=True)
        raw = F.interpolate(v.float(), size=(oh, ow), mode="bicubic", antialias=True, align_corners=False)
        assert out.dtype == torch.uint8 and torch.equal(out, raw.clamp(0, 255).round().to(torch.uint8))
        np.savez_compressed(os.path.join(HERE, f"resize_{name}.npz"), frames=v.contiguous().numpy(), hwc=hwc, out=out.numpy(),
                            raw=raw.contiguous().numpy())
        print("resize", name, tuple(out.shape), "clamped", int((raw > 255).sum() + (raw < 0).sum()))
    sys.path.insert(0, "/root/reference/train")
    import qwen_vl_utils.vision_process as vp
    rng = np.random.default_rng(0)
    sizes, frames = [], []
    for _ in range(300):
        h, w = int(rng.integers(20, 2200)), int(rng.integers(20, 2200))
        if max(h, w) / min(h, w) > 150:
            continue
        lo = int(rng.choice([4 * 28 * 28, 128 * 28 * 28, 256 * 28 * 28]))
        hi = int(rng.choice([192 * 28 * 28, 151200, 768 * 28 * 28, 16384 * 28 * 28]))
        if hi < lo:
            continue
        sizes.append([h, w, lo, hi, *vp.smart_resize(h, w, 28, lo, hi)])
    for h, w in ((720, 1280), (1080, 1920), (360, 640), (480, 640), (14, 14), (28, 5000), (1000, 1014), (42, 70)):
        sizes.append([h, w, vp.VIDEO_MIN_PIXELS, vp.VIDEO_MAX_PIXELS, *vp.smart_resize(h, w, 28, vp.VIDEO_MIN_PIXELS, vp.VIDEO_MAX_PIXELS)])
    for _ in range(200):
        total, vfps = int(rng.integers(2, 20000)), float(rng.choice([23.976, 24.0, 25.0, 29.97, 30.0, 60.0, 5.0]))
        fps, mx = float(rng.choice([2.0, 1.0, 0.5, 4.0])), int(rng.choice([32, 128, 768, 7]))
        try:
            n = vp.smart_nframes({"fps": fps, "max_frames": min(mx, total)}, total, vfps)
        except ValueError:
            n = -1
        frames.append([total, vfps, fps, min(mx, total), n])
    for total, want in ((100, 7), (100, 8), (100, 9), (10, 10), (10, 12), (3, 2), (50, 1)):
        try:
            n = vp.smart_nframes({"nframes": want}, total, 30.0)
        except ValueError:
            n = -1
        frames.append([total, 30.0, -1.0, want, n])
    np.savez_compressed(os.path.join(HERE, "resize_helpers.npz"), sizes=np.asarray(sizes, dtype=np.int64),
                        frames=np.asarray(frames, dtype=np.float64), fps_max_frames=vp.FPS_MAX_FRAMES,
                        total_pixels=vp.VIDEO_TOTAL_PIXELS)
    print("helpers", len(sizes), len(frames))


def rope_goldens():
    """Pins for oracle/rope_index.py (the transformers-4.51.1 get_rope_index restatement) from the two live
    implementations that exist in this image (4.51.1 itself is not installed and there is no network):

      rope_vllm.json   every multimodal row of tests/rope_cases.py::case_suite (unpadded) + extra exactly-representable
                       second_per_grid_ts cases through the INSTALLED vLLM 0.22 `Qwen2_5_VLForConditionalGeneration.
                       get_mrope_input_positions` (+ its `iter_mm_grid_thw`), called unbound on a stub `self` that
                       carries only `config.vision_config`.  vLLM multiplies in float64: it pins the
                       multiply-then-truncate mode wherever fp32 and fp64 truncation agree (always for spg values
                       that are exact in binary: 0.5, 1, 2, 2.5, 9.375, ...).
      rope_hf55.json   rows whose vision items all have t == 1 through the INSTALLED transformers 5.5.0
                       `Qwen2_5_VLModel.get_rope_index` (stub self, tokens_per_second = 1).  5.5.0 diverges from 4.51.1
                       for t > 1 (constant temporal index, different h order: SURVEY §8c), zero-fills padding and uses
                       the unpadded length in the delta; for t == 1 items, unmasked positions and
                       `delta + (S - unpadded length)` coincide with 4.51.1 and are what the tests compare.
    """
    import json
    import types
    sys.path.insert(0, os.path.join(HERE, ".."))
    from rope_cases import IMG, VID, VS, build_row, case_suite, pad_batch
    from vllm.model_executor.models.qwen2_5_vl import Qwen2_5_VLForConditionalGeneration as VL
    NS = types.SimpleNamespace

    def vllm_row(tokens, img, vid, spg, tps):
        stub = NS(config=NS(vision_config=NS(spatial_merge_size=2, tokens_per_second=tps)))
        stub.iter_mm_grid_thw = types.MethodType(VL.iter_mm_grid_thw, stub)
        feats, ii, vi = [], 0, 0
        for j, tok in enumerate(tokens[:-1]):
            if tok != VS:
                continue
            if tokens[j + 1] == IMG:
                feats.append(NS(mm_position=NS(offset=j + 1), modality="image",
                                data={"image_grid_thw": NS(data=torch.tensor(img[ii]))}))
                ii += 1
            elif tokens[j + 1] == VID:
                d = {"video_grid_thw": NS(data=torch.tensor(vid[vi]))}
                if spg is not None:
                    d["second_per_grid_ts"] = NS(data=torch.tensor(spg[vi], dtype=torch.float32))
                feats.append(NS(mm_position=NS(offset=j + 1), modality="video", data=d))
                vi += 1
        pos, delta = VL.get_mrope_input_positions(stub, list(tokens), feats)
        return pos.tolist(), int(delta)

    rows = []

    def add_rows(name, kw, tps=2):
        ids, mask = kw["input_ids"], kw.get("attention_mask")
        spg = kw.get("second_per_grid_ts")
        spg32 = None if spg is None else torch.as_tensor(spg, dtype=torch.float32)
        ii = vi = 0
        for b in range(ids.shape[0]):
            toks = (ids[b][mask[b].bool()] if mask is not None else ids[b]).tolist()
            n_i = sum(1 for j, v in enumerate(toks[:-1]) if v == VS and toks[j + 1] == IMG)
            n_v = sum(1 for j, v in enumerate(toks[:-1]) if v == VS and toks[j + 1] == VID)
            if n_i + n_v == 0:
                continue
            img = kw["image_grid_thw"][ii:ii + n_i].tolist() if n_i else []
            vid = kw["video_grid_thw"][vi:vi + n_v].tolist() if n_v else []
            s = None if spg32 is None or not n_v else [float(spg32[vi + k]) for k in range(n_v)]
            ii, vi = ii + n_i, vi + n_v
            pos, delta = vllm_row(toks, img, vid, s, tps)
            rows.append({"name": f"{name}[{b}]", "tokens": toks, "image_grid_thw": img, "video_grid_thw": vid,
                         "second_per_grid_ts": s, "tokens_per_second": tps, "pos": pos, "delta": delta})

    for name, kw in case_suite():
        if "image_grid_thw" in kw or "video_grid_thw" in kw:
            add_rows(name, kw)
    rng = np.random.default_rng(7)
    # exactly representable seconds-per-grid (fp32 == fp64 for every k): the whole temporal row is pinned
    for spg in (0.5, 0.25, 1.5, 2.5, 9.375, 4.0, 0.125):
        g = [int(rng.integers(2, 14)), 4, 6]
        add_rows(f"exact_spg_{spg}", dict(input_ids=torch.tensor([build_row([(VID, g)], rng)]),
                                          video_grid_thw=torch.tensor([g]), second_per_grid_ts=torch.tensor([spg])))
    # the HF docstring shape: tokens_per_second 25, 2 s per grid
    add_rows("docstring_tps25", dict(input_ids=torch.tensor([[VS] + [VID] * 12 + [11, 12, 13, 14, 15]]),
                                     video_grid_thw=torch.tensor([[3, 4, 4]]), second_per_grid_ts=torch.tensor([2.0])), tps=25)
    # BASELINE config 2 sequence (32 frames 448x448), fps 2 -> 1 s per grid
    g = [16, 32, 32]
    add_rows("config2", dict(input_ids=torch.tensor([build_row([(VID, g)], rng, lead=32, gap=0, tail=64)]),
                             video_grid_thw=torch.tensor([g]), second_per_grid_ts=torch.tensor([1.0])))
    with open(os.path.join(HERE, "rope_vllm.json"), "w") as f:
        json.dump({"source": "vllm 0.22 Qwen2_5_VLForConditionalGeneration.get_mrope_input_positions (float64 temporal index)",
                   "rows": rows}, f, separators=(",", ":"))
    print("rope vllm rows", len(rows))

    # ---- transformers 5.5.0, t == 1 items only
    from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as M
    hf_stub = NS(config=NS(vision_config=NS(spatial_merge_size=2, tokens_per_second=1)))
    hf_stub.get_vision_position_ids = types.MethodType(M.Qwen2_5_VLModel.get_vision_position_ids, hf_stub)
    batches = []
    im1, im2, im3, v1 = [1, 6, 4], [1, 2, 2], [1, 10, 8], [1, 4, 12]
    specs = {
        "images_right": ([[(IMG, im1), (IMG, im2)], [(IMG, im3)], []], "right"),
        "images_left": ([[(IMG, im1), (IMG, im2)], [(IMG, im3)], []], "left"),
        "t1_video_and_image": ([[(VID, v1), (IMG, im2)], [(IMG, im1), (VID, v1)]], "left"),
    }
    for name, (spec, side) in specs.items():
        rws = [build_row(items, rng) if items else rng.integers(1000, 5000, size=9).tolist() for items in spec]
        ids, mask = pad_batch(rws, side)
        img = [g for items in spec for k, g in items if k == IMG]
        vid = [g for items in spec for k, g in items if k == VID]
        types_ = torch.zeros_like(ids)
        types_[ids == IMG] = 1
        types_[ids == VID] = 2
        pos, delta = M.Qwen2_5_VLModel.get_rope_index(
            hf_stub, ids, types_, torch.tensor(img) if img else None, torch.tensor(vid) if vid else None, None, mask)
        batches.append({"name": name, "input_ids": ids.tolist(), "attention_mask": mask.tolist(), "image_grid_thw": img,
                        "video_grid_thw": vid, "tokens_per_second": 1, "pos": pos.tolist(), "delta": delta.flatten().tolist()})
    import transformers
    with open(os.path.join(HERE, "rope_hf55.json"), "w") as f:
        json.dump({"source": f"transformers {transformers.__version__} Qwen2_5_VLModel.get_rope_index, t == 1 items, tokens_per_second 1",
                   "batches": batches}, f, separators=(",", ":"))
    print("rope hf55 batches", len(batches))


if __name__ == "__main__":
    which = sys.argv[1:] or ["tower", "patchify", "mrope", "resize", "rope"]
    if "rope" in which:
        rope_goldens()
    if "tower" in which:
        main()
    if "patchify" in which:
        patchify_goldens()
    if "mrope" in which:
        mrope_goldens()
    if "resize" in which:
        resize_goldens()
