"""A/B timing of kernel kinds with CUDA events inside the full config-2 step (VR2_LIB selects the build)."""
import os, sys, torch
sys.path.insert(0, ".")
import video_r2_b200 as vr2
from bench import synth_state_dict
kinds = os.environ.get("KINDS", "attn_window,attn_full").split(",")
cfg = vr2.VisionConfig()
dev = torch.device("cuda:0")
tower = vr2.VisionTower(cfg, synth_state_dict(cfg, dev), device=dev, precision=os.environ.get("VR2_PRECISION", "f16"))
grid = [[16, 32, 32]]
x = torch.randn(16384, 1176, device=dev).to(torch.bfloat16)
for _ in range(5):
    tower(x, grid)
torch.cuda.synchronize()
tower.profile(kinds)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 15
for _ in range(n):
    tower(x, grid)
e1.record()
torch.cuda.synchronize()
r = tower.profile_read()
print(os.environ.get("VR2_LIB", "default"), "step %.3f ms" % (e0.elapsed_time(e1) / n),
      {k: "%.4f ms/step (%.1f us/launch)" % (v[0] / n, 1e3 * v[0] / v[1]) for k, v in r.items()})
