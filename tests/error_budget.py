"""Where the operand-rounding error of the bf16 / f16 precisions comes from: the fp32 oracle with one
rounding site of the GPU path switched on at a time (qkv input, q/k/v, probabilities, attention output, MLP input,
h, merger input, merger hidden), depth 32, grid [[2,16,16]].  CPU only, ~2 minutes:

    python tests/error_budget.py

Analysis tool, not collected by pytest."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import tower as T, indices
torch.set_num_threads(8)
SITES = ["xb_qkv", "qkv_out", "P", "attn_out", "xb_mlp", "h", "merger_in", "merger_mid"]
MODE={'dt':torch.bfloat16}
def bf(x): return x.to(MODE['dt']).to(torch.float32)

@torch.no_grad()
def fwd(weights, cfg, pixels, grid_thw, on, layers_split=()):
    r = lambda name, x, l=None: bf(x) if (name in on and (l is None or l not in layers_split)) else x
    W = {k: v.to(torch.float32) for k, v in weights.items()}
    x = pixels.to(torch.float32)
    grid = np.asarray(grid_thw, dtype=np.int64).reshape(-1, 3)
    h, H, D = cfg.hidden_size, cfg.num_heads, cfg.head_dim
    unit = 4
    x = x @ W["patch_embed.proj.weight"].reshape(h, -1).T
    win, cu_win, _ = indices.window_index(grid, 2, cfg.window_size, cfg.patch_size)
    cu_full = indices.cu_seqlens(grid)
    cos, sin = indices.rotary_tables(grid, win, D, cfg.rope_theta, 2)
    cos = torch.from_numpy(np.concatenate([cos, cos], axis=-1))[:, None, :]
    sin = torch.from_numpy(np.concatenate([sin, sin], axis=-1))[:, None, :]
    N = x.shape[0]
    x = x.reshape(N // unit, unit, h)[torch.from_numpy(win)].reshape(N, h)
    scale = 1 / math.sqrt(D)
    for l in range(cfg.depth):
        p = f"blocks.{l}."
        cu = cu_full if l in cfg.fullatt_block_indexes else cu_win
        rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + cfg.rms_norm_eps)
        y = r("xb_qkv", x, l) * rstd * W[p + "norm1.weight"]
        qkv = F.linear(y, W[p + "attn.qkv.weight"], W[p + "attn.qkv.bias"]).reshape(N, 3, H, D)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        q = r("qkv_out", q * cos + T._rotate_half(q) * sin, l)
        k = r("qkv_out", k * cos + T._rotate_half(k) * sin, l)
        v = r("qkv_out", v, l)
        # attention with optional P rounding
        out = torch.empty_like(q)
        cua = np.asarray(cu); lens = cua[1:] - cua[:-1]
        for L in np.unique(lens):
            if L == 0: continue
            starts = cua[:-1][lens == L]
            idx = torch.from_numpy((starts[:, None] + np.arange(L)[None, :]).reshape(-1))
            qs = q[idx].reshape(len(starts), L, H, D).transpose(1, 2); ks = k[idx].reshape(len(starts), L, H, D).transpose(1, 2); vs = v[idx].reshape(len(starts), L, H, D).transpose(1, 2)
            s = torch.matmul(qs, ks.transpose(-1, -2)) * scale
            m = s.max(-1, keepdim=True)[0]
            pexp = torch.exp(s - m)
            den = pexp.sum(-1, keepdim=True)
            o = torch.matmul(r("P", pexp, l), vs) / den
            out[idx] = o.transpose(1, 2).reshape(-1, H, D)
        a = r("attn_out", out.reshape(N, h), l)
        x = x + F.linear(a, W[p + "attn.proj.weight"], W[p + "attn.proj.bias"])
        rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + cfg.rms_norm_eps)
        y = r("xb_mlp", x, l) * rstd * W[p + "norm2.weight"]
        g = F.linear(y, W[p + "mlp.gate_proj.weight"], W[p + "mlp.gate_proj.bias"])
        u = F.linear(y, W[p + "mlp.up_proj.weight"], W[p + "mlp.up_proj.bias"])
        x = x + F.linear(r("h", F.silu(g) * u, l), W[p + "mlp.down_proj.weight"], W[p + "mlp.down_proj.bias"])
    rstd = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + cfg.rms_norm_eps)
    y = (r("merger_in", x) * rstd * W["merger.ln_q.weight"]).reshape(N // unit, unit * h)
    y = r("merger_mid", F.gelu(F.linear(y, W["merger.mlp.0.weight"], W["merger.mlp.0.bias"])))
    y = F.linear(y, W["merger.mlp.2.weight"], W["merger.mlp.2.bias"])
    return y[torch.from_numpy(indices.reverse_index(win))], x

cfg = T.TowerConfig()
w = T.make_weights(cfg, seed=0)
grid = [[2, 16, 16]]
pix = bf(torch.randn(512, 1176, generator=torch.Generator().manual_seed(1)))
ref, xref = fwd(w, cfg, pix, grid, set())
print("ref |out| max", ref.abs().max().item(), "rms", ref.pow(2).mean().sqrt().item(), "|x| max", xref.abs().max().item(), "rms", xref.pow(2).mean().sqrt().item())
def report(name, on, ls=()):
    o, x = fwd(w, cfg, pix, grid, set(on), ls)
    e = (o - ref).abs().max().item(); c = F.cosine_similarity(o.flatten(), ref.flatten(), dim=0).item()
    ex = (x - xref).abs().max().item()
    print(f"{name:40s} out max-abs {e:.4f} cos {c:.7f}   x err {ex:.4f}")

report("bf16 all", SITES)
MODE['dt']=torch.float16
report("fp16 all", SITES)
for s_ in SITES: report("fp16 only " + s_, [s_])
