"""Profiling driver: a depth-8 tower (layers 0-6 window attention, layer 7 full attention) on the config-2 clip
([16,32,32]); two forwards.  Used under ncu with -k regex:<kernel> -s <launches of the first forward> (see profiles/README.md)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import video_r2_b200 as vr2  # noqa: E402
from bench import synth_state_dict  # noqa: E402

depth = int(os.environ.get("VR2_PROF_DEPTH", "8"))
grid = [[int(v) for v in os.environ.get("VR2_PROF_GRID", "16,32,32").split(",")]]
cfg = vr2.VisionConfig(depth=depth, fullatt_block_indexes=(depth - 1,))
dev = torch.device("cuda:0")
tower = vr2.VisionTower(cfg, synth_state_dict(cfg, dev), device=dev, precision=os.environ.get("VR2_PRECISION", "f16"))
n = grid[0][0] * grid[0][1] * grid[0][2]
x = torch.randn(n, cfg.patch_dim, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).to(torch.bfloat16)
for _ in range(int(os.environ.get("VR2_PROF_ITERS", "2"))):
    out = tower(x, grid)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.float().abs().max()))
