"""CPU: the M-RoPE oracle against golden outputs of HF's own rotary code (tests/golden/mrope_small.npz, written by
tests/golden/make_golden.py) and against the live HF functions when importable."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import mrope as O


@pytest.mark.parametrize("name,dt", [("f32", torch.float32), ("bf16", torch.bfloat16)])
def test_oracle_matches_hf_golden(name, dt):
    g = np.load(os.path.join(GOLDEN, "mrope_small.npz"))
    q, k = torch.from_numpy(g[f"q_{name}"]).to(dt), torch.from_numpy(g[f"k_{name}"]).to(dt)
    rq, rk = O.apply_mrope(q, k, torch.from_numpy(g["pos"]), torch.from_numpy(g["inv_freq"]), g["section"].tolist())
    assert torch.equal(rq.float(), torch.from_numpy(g[f"rq_{name}"])) and torch.equal(rk.float(), torch.from_numpy(g[f"rk_{name}"]))


def test_oracle_matches_live_hf_functions():
    try:
        from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as M
    except Exception as e:
        pytest.skip(f"transformers unavailable: {e!r}")
    g = torch.Generator().manual_seed(2)
    B, S, D = 3, 50, 128
    pos = torch.randint(0, 20000, (3, B, S), generator=g)
    inv = 1.0 / (1000000.0 ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    for dt in (torch.float32, torch.bfloat16):
        q, k = torch.randn(B, 5, S, D, generator=g).to(dt), torch.randn(B, 1, S, D, generator=g).to(dt)
        freqs = (inv[None, None, :, None].float().expand(3, B, -1, 1) @ pos[:, :, None, :].float()).transpose(2, 3)
        emb = torch.cat((freqs, freqs), dim=-1)
        rq, rk = M.apply_multimodal_rotary_pos_emb(q, k, emb.cos().to(dt), emb.sin().to(dt), [16, 24, 24])
        oq, ok = O.apply_mrope(q, k, pos, inv)
        assert torch.equal(rq, oq) and torch.equal(rk, ok)
