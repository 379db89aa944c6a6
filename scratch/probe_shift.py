import sys, torch
sys.path.insert(0, ".")
import video_r2_b200 as vr2
from video_r2_b200 import ops
from oracle import tower as T
# 1. does the tensor core honour fp16 subnormal inputs?
a = torch.full((128, 64), 1e-5, dtype=torch.float16, device="cuda")      # subnormal in fp16 (min normal 6.1e-5)
w = torch.ones((32, 64), dtype=torch.float16, device="cuda")
out = ops.linear(a, w)
print("subnormal A x ones: got", out[0, 0].item(), "expected", 64 * float(a[0, 0]))
w2 = torch.full((32, 64), 1e-5, dtype=torch.float16, device="cuda")
a2 = torch.ones((128, 64), dtype=torch.float16, device="cuda")
print("ones x subnormal W: got", ops.linear(a2, w2)[0, 0].item(), "expected", 64 * float(w2[0, 0]))
# 2. which shift makes the difference on in-range data?
cfg = T.TowerConfig(depth=4, fullatt_block_indexes=(1, 3))
wts = T.make_weights(cfg, seed=21)
GRID = [[2, 8, 12], [1, 4, 6]]
N = 216
x = torch.randn(N, 1176, generator=torch.Generator().manual_seed(22)).to(torch.bfloat16).cuda()
ref = T.tower_forward(wts, cfg, x.float().cpu(), GRID)
g = vr2.VisionTower(vr2.VisionConfig(depth=4, fullatt_block_indexes=(1, 3)), wts, device="cuda:0", precision="f16")
plain = g(x, GRID, out_dtype=torch.float32).clone()
print("plain vs oracle", (plain.cpu() - ref).abs().max().item())
for sx, sh in ((1, 0), (4, 0), (0, 1), (0, 4), (4, 4), (8, 8), (0, 0)):
    g.set_f16_shift(sx, sh)
    o = g(x, GRID, out_dtype=torch.float32)
    print((sx, sh), "vs plain", (o - plain).abs().max().item(), "vs oracle", (o.cpu() - ref).abs().max().item(), "sat", g.saturation_counts())
