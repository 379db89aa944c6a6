"""GPU: every fused stage of a transformer block against a plain fp32 PyTorch evaluation of the SAME stage, fed with
what the previous stage of the kernel path actually wrote.

The golden-vector tests compare whole towers, where an error in one fused epilogue hides behind the operand rounding of
32 layers.  Here the forward is stopped after each stage (vr2_tower_set_debug_stop) and the workspace is read back
(vr2_tower_workspace_layout), so each GEMM + epilogue — patch embed with the window permute, QKV with the folded RMSNorm,
bias and 2-D rotary, attention, proj / down with the fp32 residual, the next layer's operand and its sum of squares,
gate|up with the SwiGLU product — is checked on its own, to one rounding of its 16-bit output, on a grid with partial
windows and a row count that fills no tile (HF modeling_qwen2_5_vl.py:57-88, :108-114, :149-167, :221-321)."""
import math

import numpy as np
import pytest
import torch

from oracle import tower as T

pytestmark = pytest.mark.gpu

GRID = [[2, 26, 18], [3, 16, 16]]     # merged 13 x 9 (1- and 3-wide partial windows) + 8 x 8; 1704 patch rows


def _segments(x, cu):
    return [(int(cu[i]), int(cu[i + 1])) for i in range(len(cu) - 1) if cu[i + 1] > cu[i]]


def _attention_ref(qkv, cu, heads=16, d=80):
    n = qkv.shape[0]
    q, k, v = qkv.float().reshape(n, 3, heads, d).unbind(1)
    out = torch.empty(n, heads, d, device=qkv.device)
    for s, e in _segments(qkv, cu):
        att = torch.einsum("qhd,khd->hqk", q[s:e], k[s:e]) / math.sqrt(d)
        out[s:e] = torch.einsum("hqk,khd->qhd", att.softmax(-1), v[s:e])
    return out.reshape(n, heads * d)


@pytest.mark.parametrize("prec", ["bf16", "f16"])
def test_every_stage_of_two_blocks(prec):
    import video_r2_b200 as vr2
    cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    W = T.make_weights(cfg_o, seed=21)
    tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), W, device="cuda:0", precision=prec)
    dev = torch.device("cuda:0")
    Wd = {k: v.to(dev) for k, v in W.items()}
    act = torch.float16 if prec == "f16" else torch.bfloat16
    ulp = 2.0 ** -11 if prec == "f16" else 2.0 ** -8         # one rounding to the 16-bit format, relative
    n = sum(t * h * w for t, h, w in GRID)
    hd, inter, eps = 1280, 3420, 1e-6
    pix = torch.randn(n, 1176, generator=torch.Generator().manual_seed(22)).to(torch.bfloat16).to(dev)
    plan = tower.plan(GRID)
    ex = plan.export()
    cos, sin = plan.export_rotary()                             # [N, 40] fp32, window order
    cu_win, cu_full = ex["cu_window_seqlens"].cpu().numpy(), ex["cu_seqlens"].cpu().numpy()
    win = ex["window_index"]

    def run_to(stage):
        tower.debug_stop(stage)
        try:
            tower(pix, GRID)
            torch.cuda.synchronize()
            return {k: v.clone() for k, v in tower.debug_buffers(n).items()}
        finally:
            tower.debug_stop(0)

    def close16(got, ref, what):      # a 16-bit buffer against its fp32 reference: one rounding (+ accumulation order)
        got, ref = got.float(), ref.float()
        tol = ulp * ref.abs() + 2e-3 * ulp * 256 + 1e-4
        bad = (got - ref).abs() > tol
        assert not bad.any(), (what, prec, int(bad.sum()), float((got - ref).abs().max()))

    def check_resid(buf, x_ref, norm_w, parts, what):
        # fp32 residual stream: exact bf16 / fp16 products accumulated in fp32 — only the summation order differs
        assert (buf["x"] - x_ref).abs().max().item() < 2e-3, what
        # the next GEMM's operand is ONE rounding of the kernel's own x times the norm weight: the same fp32 product
        xb_ref = (buf["x"] * norm_w).to(act)
        mism = (buf["xb"] != xb_ref).float().mean().item()
        assert mism < 1e-4, (what, "xb", mism)     # (a fused multiply-add may differ from mul-then-round in the last bit)
        ssq = buf["ssq"][:parts].sum(0)
        ref = buf["x"].pow(2).sum(-1)
        assert ((ssq - ref).abs() <= 1e-4 * ref + 1e-6).all(), (what, "ssq")

    # ---- stage 1: patch embed (Conv3d as a GEMM, no bias) stored through the window permute
    s1 = run_to(1)
    x0 = pix.float() @ Wd["patch_embed.proj.weight"].reshape(hd, -1).T
    x0 = x0.reshape(n // 4, 4, hd)[win].reshape(n, hd)
    check_resid(s1, x0, Wd["blocks.0.norm1.weight"], hd // 40, "patch_embed")
    prev, stage, parts = s1, 1, hd // 40

    for l in range(2):
        p = f"blocks.{l}."
        full = l == 1
        # ---- qkv: rstd (x (.) w) W^T + b, 2-D rotary on the q and k heads
        cur = run_to(stage + 1)
        rstd = torch.rsqrt(prev["ssq"][:parts].sum(0) / hd + eps)
        y = (prev["xb"].float() @ Wd[p + "attn.qkv.weight"].T) * rstd[:, None] + Wd[p + "attn.qkv.bias"]
        y = y.reshape(n, 3, 16, 80)
        c = torch.cat([cos, cos], -1)[:, None, :]
        s_ = torch.cat([sin, sin], -1)[:, None, :]
        rot = lambda t: t * c + torch.cat((-t[..., 40:], t[..., :40]), -1) * s_
        ref = torch.stack([rot(y[:, 0]), rot(y[:, 1]), y[:, 2]], 1).reshape(n, 3 * hd)
        close16(cur["qkv"], ref, f"qkv[{l}]")
        # ---- attention over the windows (block 0) / the temporal slices (block 1), from the kernel's own q | k | v
        att = run_to(stage + 2)
        ref = _attention_ref(cur["qkv"], cu_full if full else cu_win)
        err = (att["att"].float() - ref).abs().max().item()
        assert err < (2e-2 if prec == "bf16" else 4e-3), (f"attention[{l}]", err)
        # ---- proj + residual, operand and sum of squares for norm2
        pj = run_to(stage + 3)
        x1 = att["x"] + att["att"].float() @ Wd[p + "attn.proj.weight"].T + Wd[p + "attn.proj.bias"]
        check_resid(pj, x1, Wd[p + "norm2.weight"], 5, f"proj[{l}]")
        # ---- gate | up with the SwiGLU product
        gu = run_to(stage + 4)
        rstd = torch.rsqrt(pj["ssq"][:5].sum(0) / hd + eps)
        xb = pj["xb"].float()
        g = (xb @ Wd[p + "mlp.gate_proj.weight"].T) * rstd[:, None] + Wd[p + "mlp.gate_proj.bias"]
        u = (xb @ Wd[p + "mlp.up_proj.weight"].T) * rstd[:, None] + Wd[p + "mlp.up_proj.bias"]
        close16(gu["h"][:, :inter], torch.nn.functional.silu(g) * u, f"swiglu[{l}]")
        assert (gu["h"][:, inter:].float() == 0).all(), "the K padding of h must be zero (down multiplies it)"
        # ---- down + residual, operand and sum of squares for the next block's norm1 (the merger's ln_q after the last)
        dn = run_to(stage + 5)
        x2 = gu["x"] + gu["h"][:, :inter].float() @ Wd[p + "mlp.down_proj.weight"].T + Wd[p + "mlp.down_proj.bias"]
        nxt = Wd["blocks.1.norm1.weight"] if l == 0 else Wd["merger.ln_q.weight"]
        check_resid(dn, x2, nxt, 5, f"down[{l}]")
        prev, stage, parts = dn, stage + 5, 5

    # ---- PatchMerger on the kernel's own final residual stream: ln_q -> 16 bit, fc1 + GELU(erf) -> 16 bit, fc2 in fp32,
    #      rows back in the original token order (HF :133-146, :512-513)
    out = tower(pix, GRID, out_dtype=torch.float32)
    x = prev["x"]
    y = (Wd["merger.ln_q.weight"] * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))).to(act).float()
    hid = torch.nn.functional.gelu(y.reshape(n // 4, 4 * hd) @ Wd["merger.mlp.0.weight"].T + Wd["merger.mlp.0.bias"])
    ref = hid.to(act).float() @ Wd["merger.mlp.2.weight"].T + Wd["merger.mlp.2.bias"]
    ref = ref[ex["reverse_index"]]
    err = (out - ref).abs().max().item()
    # the two 16-bit roundings are reproduced above; what is left are last-bit flips of them and the summation order
    assert err < (2e-2 if prec == "bf16" else 3e-3), ("merger", prec, err)
