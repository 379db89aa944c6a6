"""The library bar: Hugging Face's own Qwen2.5-VL vision tower in bf16 on the same B200 — cuBLAS GEMMs,
flash-attn 2 or sdpa attention, eager elementwise — timed on BASELINE configs[1] beside this repo's tower.

    python profiles/hf_gpu_bar.py [--steps 10] [--warmup 3] > profiles/rNN_hf_gpu_bar.json

Measurement only: nothing in the product imports this file.  transformers ships in the image (no /root/reference
needed); weights are the same random init bench.py uses."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402

CLIP = bench.CLIP


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = CLIP[0] * CLIP[1] * CLIP[2]
    tokens = n // 4
    grid = [list(CLIP)]
    x = torch.randn(n, 1176, generator=torch.Generator(device=dev).manual_seed(1), device=dev).to(torch.bfloat16)
    out = {"workload": f"configs[1] grid_thw {grid}, {tokens} visual tokens, bf16, 1 B200", "results": {}}

    model = bench.build_hf_reference().to(dev).to(torch.bfloat16)
    gthw = torch.tensor(grid, device=dev)
    ref_out = None
    for impl in ("sdpa", "flash_attention_2"):
        try:
            model.config._attn_implementation = impl

            def run():
                with torch.no_grad():
                    r = model(x, grid_thw=gthw)
                return r.pooler_output if hasattr(r, "pooler_output") else r

            ref_out = run()
            ms = timed(run, args.steps, args.warmup)
            out["results"]["hf_" + impl] = {"ms": ms, "tokens_per_s": tokens / ms * 1e3}
        except Exception as e:
            out["results"]["hf_" + impl] = {"error": repr(e)[:300]}

    import video_r2_b200 as vr2
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    tower = vr2.VisionTower(vr2.VisionConfig(), sd, device=dev)
    if True:
        plan = tower.plan(grid)
        ms = timed(lambda: tower.forward(x, plan), args.steps, args.warmup)
        out["results"]["vr2_bf16"] = {"ms": ms, "tokens_per_s": tokens / ms * 1e3}
        if ref_out is not None:
            mine = tower.forward(x, plan).float()
            out["results"]["vr2_bf16"]["max_abs_vs_hf_bf16"] = (mine - ref_out.float()).abs().max().item()
        for k, v in list(out["results"].items()):
            if k.startswith("hf_") and "ms" in v:
                v["vr2_speedup"] = v["ms"] / ms
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
