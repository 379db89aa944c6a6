"""First-light diagnostic (not a pytest file): prints how the tcgen05 GEMM compares with torch,
with enough structure to tell a descriptor / swizzle bug from a pipeline bug."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_r2_b200 import ops

def run(M, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    out = ops.linear(a, w, None)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T
    err = (out - ref).abs()
    print(f"M={M} N={N} K={K}: max err {err.max().item():.4g}  ref absmax {ref.abs().max().item():.3g}", flush=True)
    if err.max().item() > 1e-2:
        bad = err > 1e-2
        print("  bad fraction", bad.float().mean().item())
        print("  bad rows (first 16 of tile):", bad[:16].any(1).int().tolist())
        print("  bad cols per 8-block:", bad[:128, :N].reshape(min(M,128), -1, 8).any(-1).any(0).int().tolist()[:32])
        # is it a K-slice problem? compare against partial-K references
        for kk in range(16, min(K, 128) + 1, 16):
            pr = a[:, :kk].float() @ w[:, :kk].float().T
            print(f"   vs first {kk} of K: max err {(out - pr).abs().max().item():.4g}")
        print("  out[0,:8]", out[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
    return err.max().item()

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    for shape in [(128, 256, 64), (128, 256, 128), (256, 512, 256), (300, 1280, 1176), (8192, 3840, 1280)]:
        run(*shape)
    # timing of the big shapes
    for (M, N, K) in [(16384, 3840, 1280), (16384, 1280, 1280), (16384, 1280, 3456), (16384, 5120, 1280), (4096, 5120, 5120)]:
        a = torch.randn(M, K).to(torch.bfloat16).cuda(); w = torch.randn(N, K).to(torch.bfloat16).cuda()
        for _ in range(3): ops.linear(a, w, None, out_dtype=torch.bfloat16)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.linear(a, w, None, out_dtype=torch.bfloat16)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t0 = time.time()
        for _ in range(3): torch.nn.functional.linear(a, w)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): torch.nn.functional.linear(a, w)
        e1.record(); torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        print(f"GEMM {M}x{N}x{K}: vr2 {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s | cuBLAS {ms_t:.3f} ms = {2*M*N*K/ms_t/1e9:.0f} TF/s", flush=True)
