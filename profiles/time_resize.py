"""Timing of vr2.resize_frames against torch's CPU kernel on the box.  Measurement only."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, time, json
import torch.nn.functional as F
import video_r2_b200 as vr2
res = {}
for (T, H, W, mp) in ((32, 720, 1280, None), (128, 720, 1280, 151200), (32, 1080, 1920, None)):
    x = torch.randint(0, 256, (T, H, W, 3), dtype=torch.uint8).permute(0, 3, 1, 2)
    size = vr2.video_target_size(T, H, W, max_pixels=mp)
    xd = x.cuda()
    xc = xd.contiguous()
    for name, src in (("hwc", xd), ("chw", xc)):
        for _ in range(3): out = vr2.resize_frames(src, size)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): out = vr2.resize_frames(src, size)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        byts = T * 3 * (H * W + size[0] * size[1])
        res[f"{T}x{H}x{W}->{size} {name}"] = {"ms": ms, "GBps": byts / ms / 1e6}
    torch.set_num_threads(torch.get_num_threads())
    t0 = time.perf_counter()
    ref = F.interpolate(x.float(), size=size, mode="bicubic", antialias=True, align_corners=False).clamp(0, 255).round().to(torch.uint8)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    res[f"{T}x{H}x{W}->{size} cpu"] = {"ms": cpu_ms, "threads": torch.get_num_threads(), "equal": bool(torch.equal(out.cpu(), ref))}
print(json.dumps(res, indent=1))
