import sys, torch
sys.path.insert(0, ".")
import video_r2_b200 as vr2
from oracle import tower as T
GRID = [[2, 8, 12], [1, 4, 6]]
N = 216
x = torch.randn(N, 1176, generator=torch.Generator().manual_seed(22)).to(torch.bfloat16).cuda()
for depth in (1, 2):
    cfg = T.TowerConfig(depth=depth, fullatt_block_indexes=(0,))
    wts = T.make_weights(cfg, seed=21)
    g = vr2.VisionTower(vr2.VisionConfig(depth=depth, fullatt_block_indexes=(0,)), wts, device="cuda:0", precision="f16")
    plain = g(x, GRID, out_dtype=torch.float32).clone()
    for sx, sh in ((1, 0), (0, 1), (4, 4)):
        g.set_f16_shift(sx, sh)
        o = g(x, GRID, out_dtype=torch.float32)
        d = (o - plain).abs()
        print(depth, (sx, sh), "max diff", d.max().item(), "n differing", int((d > 0).sum()), "of", d.numel())
