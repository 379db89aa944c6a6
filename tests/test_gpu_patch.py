"""GPU: the drop-in boundary itself — after vr2.install() the Hugging Face classes Video-R2 calls
(monkey_patch_forward.py:392, :415-421) run libvr2.so and agree with the unpatched HF code."""
import types

import pytest
import torch

from conftest import hf_available
from oracle import rope_index as R
from oracle import tower as T
from rope_cases import case_suite

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not hf_available(), reason="transformers not importable")]


def _hf_visual(depth, full):
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel as V
    cfg = Qwen2_5_VLVisionConfig(depth=depth, hidden_size=1280, intermediate_size=3420, num_heads=16,
                                 out_hidden_size=3584, patch_size=14, spatial_merge_size=2, temporal_patch_size=2,
                                 window_size=112, fullatt_block_indexes=list(full), tokens_per_second=2,
                                 hidden_act="silu")
    cfg._attn_implementation = "sdpa"
    return V._from_config(cfg).eval()


def test_patched_visual_forward_matches_unpatched_hf():
    import video_r2_b200 as vr2
    m = _hf_visual(2, (1,))
    m.load_state_dict(T.make_weights(T.TowerConfig(depth=2, fullatt_block_indexes=(1,)), seed=31))
    m = m.cuda().float()
    grid = torch.tensor([[2, 10, 14], [1, 4, 6]], device="cuda")
    x = torch.randn(2 * 140 + 24, 1176, generator=torch.Generator().manual_seed(32)).to(torch.bfloat16).cuda()
    with torch.no_grad():
        ref = m(x.float(), grid_thw=grid)
    ref = ref.pooler_output if hasattr(ref, "pooler_output") else ref
    vr2.install()
    try:
        out = m(x, grid_thw=grid)          # the reference's call expression, unchanged
        out = out.pooler_output if hasattr(out, "pooler_output") else out
        assert out.dtype == torch.bfloat16 and out.shape == ref.shape
        err = (out.float() - ref).abs().max().item()
        cos = torch.nn.functional.cosine_similarity(out.float().flatten(), ref.flatten(), dim=0).item()
        assert err < 4e-2 and cos > 0.9995, (err, cos)
        # ZeRO-3 text-only dummy (monkey_patch_forward.py:362-366)
        dummy = m(torch.zeros(784, 1176, device="cuda", dtype=torch.bfloat16), grid_thw=torch.tensor([[1, 28, 28]], device="cuda"))
        dummy = dummy.pooler_output if hasattr(dummy, "pooler_output") else dummy
        assert dummy.shape == (196, 3584) and torch.isfinite(dummy.float()).all()
    finally:
        vr2.uninstall()
    # the same through install(precision="f16"): the tower is rebuilt in that mode and lands closer to the fp32 path
    vr2.refresh(m)
    vr2.install(precision="f16")
    try:
        out_h = m(x, grid_thw=grid)
        out_h = out_h.pooler_output if hasattr(out_h, "pooler_output") else out_h
        assert m.__dict__["_vr2_tower"].precision == "f16"
        assert (out_h.float() - ref).abs().max().item() < 2e-2
    finally:
        vr2.uninstall()
        vr2.install(precision=None)
        vr2.uninstall()
        vr2.refresh(m)
    with torch.no_grad():
        again = m(x.float(), grid_thw=grid)
    again = again.pooler_output if hasattr(again, "pooler_output") else again
    assert torch.equal(again, ref)  # uninstall restored the HF forward


def test_patched_get_rope_index_matches_oracle():
    import video_r2_b200 as vr2
    from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as M
    vis = types.SimpleNamespace(spatial_merge_size=2, tokens_per_second=2)
    fake = types.SimpleNamespace(config=types.SimpleNamespace(vision_config=vis, image_token_id=151655,
                                                              video_token_id=151656, vision_start_token_id=151652))
    owner = M.Qwen2_5_VLModel if "get_rope_index" in M.Qwen2_5_VLModel.__dict__ else M.Qwen2_5_VLForConditionalGeneration
    vr2.install()
    try:
        for name, kw in case_suite():
            g = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
            pos, delta = owner.get_rope_index(fake, g["input_ids"], image_grid_thw=g.get("image_grid_thw"),
                                              video_grid_thw=g.get("video_grid_thw"),
                                              second_per_grid_ts=g.get("second_per_grid_ts"),
                                              attention_mask=g.get("attention_mask"))
            rpos, rdelta = R.get_rope_index(kw["input_ids"], kw.get("image_grid_thw"), kw.get("video_grid_thw"),
                                            kw.get("second_per_grid_ts"), kw.get("attention_mask"))
            assert torch.equal(pos.cpu(), rpos) and torch.equal(delta.cpu(), rdelta), name
    finally:
        vr2.uninstall()
