#!/usr/bin/env python
"""Benchmark of the Video-R2 hot path on B200: Qwen2.5-VL-7B vision tower + get_rope_index.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl vr2|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch: every rank encodes its share of the batch's
temporal slices (N=1: the whole clip), the merged embeddings are all-gathered (N>1), and the
M-RoPE position ids of the batch are computed.  Workload at N=1 is BASELINE.json configs[1]
("Video-R2 eval setting": one 32-frame 448x448 clip, grid_thw [16,32,32], 16384 patches -> 4096
visual tokens, bf16); at N GPUs the batch is N such clips (weak scaling), sharded by slice.

Prints ONE JSON line (rank 0).  `value` = visual tokens/s of the whole job with inputs resident in
HBM; `e2e` = the same through the public API from pinned HOST buffers (H2D of the pixels and D2H
of the embeddings + position ids inside the timed region).  `--impl reference` times the
reference's own implementation (the Hugging Face module Video-R2 calls) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "visual_tokens_per_sec"
UNIT = "tokens/s"
CLIP = (16, 32, 32)          # configs[1]: 32 frames at 448x448 -> t=16 temporal slices of 32x32 patches
TEXT_BEFORE, TEXT_AFTER = 32, 64
FLOP_PER_TOKEN = 5.261e9     # BASELINE.md §3 (dense layers + attention, algorithmic)


# BASELINE.json configs beside the headline (configs[1]); grids are [t, h, w] in patches (t = frames / 2, h = H / 14)
CFG5_CLIPS = [(16, 448, 448), (32, 448, 448), (64, 448, 448), (128, 336, 588), (8, 280, 504), (32, 280, 504), (128, 280, 504),
              (16, 336, 588), (64, 336, 588), (4, 448, 448), (32, 392, 392), (64, 224, 392), (16, 504, 280), (128, 224, 224),
              (2, 392, 392), (48, 364, 644)]          # the fixed ragged list (frames, H, W)
SIDE_CONFIGS = {
    "cfg1_16f_448": [[8, 32, 32]],
    "cfg3_rollout_8x64f_448": [[32, 32, 32]] * 8,
    "cfg4_128f_336x588": [[64, 24, 42]],
    "cfg5_ragged_16_clips": [[f // 2, H // 14, W // 14] for f, H, W in CFG5_CLIPS],
}


def grid_flops(grid) -> float:
    """Algorithmic FLOPs of the tower on these clips (2MNK per GEMM, 4 L^2 D per attention sequence)."""
    from video_r2_b200.shard import slice_cost
    return float(sum(t * slice_cost(h, w) for t, h, w in grid))


def grid_tokens(grid) -> int:
    return sum(t * h * w for t, h, w in grid) // 4


def workload_desc(world: int) -> str:
    tokens = world * CLIP[0] * CLIP[1] * CLIP[2] // 4
    return (f"Video-R2 eval setting x{world}: {world} clip(s) of 32 frames 448x448 (grid_thw [16,32,32] each, "
            f"{tokens} visual tokens), Qwen2.5-VL-7B vision tower random init + get_rope_index; "
            "slices sharded over ranks, one all-gather of embeddings")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the HF module Video-R2 executes, on the host cores
# ---------------------------------------------------------------------------------------------
def build_hf_reference(attn="sdpa", device="cpu", dtype=None):
    import torch
    from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLVisionConfig
    from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import Qwen2_5_VisionTransformerPretrainedModel as V

    cfg = Qwen2_5_VLVisionConfig(depth=32, hidden_size=1280, intermediate_size=3420, num_heads=16,
                                 out_hidden_size=3584, patch_size=14, spatial_merge_size=2, temporal_patch_size=2,
                                 window_size=112, fullatt_block_indexes=[7, 15, 23, 31], tokens_per_second=2,
                                 hidden_act="silu")
    cfg._attn_implementation = attn
    with torch.device("meta"):
        m = V._from_config(cfg)
    m = m.to_empty(device=device)
    g = torch.Generator(device=device).manual_seed(0)
    for p in m.parameters():
        p.data.normal_(0, 0.02, generator=g)
    m.rotary_pos_emb.inv_freq = (1.0 / (10000.0 ** (torch.arange(0, 40, 2, dtype=torch.float) / 40))).to(device)
    if dtype is not None:
        m = m.to(dtype)
        m.rotary_pos_emb.inv_freq = m.rotary_pos_emb.inv_freq.float()
    return m.eval()


def cpu_reference_runner():
    """-> (kind, fn(slices) -> seconds for one forward over [[slices,32,32]] + rope index of that sample)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import rope_index as orope
    try:
        model = build_hf_reference()
        kind = "reference"

        def tower(x, grid):
            with torch.no_grad():
                r = model(x, grid_thw=torch.tensor(grid))
            return r.pooler_output if hasattr(r, "pooler_output") else r
    except Exception as e:  # transformers missing: time the oracle port instead
        print(f"[bench] HF reference unavailable ({e!r}); timing the oracle port", file=sys.stderr)
        from oracle import tower as otower
        ocfg = otower.TowerConfig()
        w = otower.make_weights(ocfg, seed=0, bf16_exact=False)
        kind = "port"

        def tower(x, grid):
            return otower.tower_forward(w, ocfg, x, grid)

    def run(slices: int) -> float:
        grid = [[slices, CLIP[1], CLIP[2]]]
        n = slices * CLIP[1] * CLIP[2]
        x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(1))
        ids = torch.tensor([[1000] * TEXT_BEFORE + [151652] + [151656] * (n // 4) + [151653] + [1000] * TEXT_AFTER])
        t0 = time.perf_counter()
        out = tower(x, grid)
        orope.get_rope_index(ids, None, torch.tensor(grid), torch.tensor([1.0]), None)
        dt = time.perf_counter() - t0
        assert out.shape[0] == n // 4
        return dt

    return kind, run


def run_reference(args, rank):
    import torch
    if rank != 0:
        return
    kind, run = cpu_reference_runner()
    t1 = run(1)  # also the first warm-up
    budget = float(os.environ.get("VR2_BENCH_REF_BUDGET_S", "150"))   # seconds of CPU work for the whole run
    slices = max(1, min(CLIP[0], int(budget / max(1e-3, (args.steps + args.warmup) * t1))))
    for _ in range(max(0, args.warmup - 1)):
        run(slices)
    times = [run(slices) for _ in range(args.steps)]
    dt = sum(times) / len(times)
    tokens = slices * CLIP[1] * CLIP[2] // 4
    value = tokens / dt
    cores = torch.get_num_threads()
    sample = f"{slices} of the {CLIP[0]} temporal slices of the config-2 clip per step ({tokens} tokens), fp32, sdpa"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(max(1, args.gpus)), "sample": sample,
                   "reference": "transformers Qwen2_5_VisionTransformerPretrainedModel (fp32, sdpa) on the host CPU, rank 0 only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def synth_state_dict(cfg, device, seed=0):
    """Random-init weights of the named architecture, generated on the device (bf16)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for name, shape in cfg.weight_shapes().items():
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("ln_q.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, device=device)
        else:
            t = 0.02 * torch.randn(shape, generator=g, device=device)
        sd[name] = t.to(torch.bfloat16)
    return sd


def run_vr2(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import video_r2_b200 as vr2
    from video_r2_b200.shard import all_gather_ragged, partition_slices

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    # one process per GPU: keep this rank's threads and pinned buffers on the GPU's own NUMA node
    host_cores = vr2.bind_to_gpu_node(dev) if world > 1 else []
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation: keep stdout to the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)          # creates the communicator now, inside the redirect
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    cfg = vr2.VisionConfig()
    tower = vr2.VisionTower(cfg, synth_state_dict(cfg, dev), device=dev, precision=args.precision)

    grid = [list(CLIP)] * world
    shares = partition_slices(grid, world)
    me = shares[rank]
    my_grid = [list(r) for r in me.grid]
    n_local = me.patch_hi - me.patch_lo
    tok_local = me.token_hi - me.token_lo
    counts = [s.token_hi - s.token_lo for s in shares]
    tokens_total = sum(counts)
    n_rot = 4  # 4 x 38.5 MB of pixels > 126 MB L2; weights (1.35 GB) + activations (0.44 GB) stream every step anyway
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    pix_dev = [torch.randn(n_local, cfg.patch_dim, generator=g, device=dev).to(torch.bfloat16) for _ in range(n_rot)]
    pix_host = [p.cpu().pin_memory() for p in pix_dev]
    # batch of `world` rows: text, <vision_start>, video pads, <vision_end>, text
    per_clip = CLIP[0] * CLIP[1] * CLIP[2] // 4
    row = [1000 + i for i in range(TEXT_BEFORE)] + [151652] + [151656] * per_clip + [151653] + [2000 + i for i in range(TEXT_AFTER)]
    ids_host = torch.tensor([row] * world, dtype=torch.int64).pin_memory()
    ids_dev = ids_host.to(dev)
    vid_thw = torch.tensor(grid, dtype=torch.int64, device=dev)
    spg = torch.tensor([2.0 / 0.7] * world, dtype=torch.float32, device=dev)
    out_host = torch.empty((tok_local, cfg.out_hidden_size), dtype=torch.bfloat16).pin_memory()
    pos_host = torch.empty((3, world, len(row)), dtype=torch.int64).pin_memory()

    # N > 1: how the merged embeddings reach every rank.  "symmetric" (default): the merger's last GEMM stores each
    # row into every rank's buffer over NVLink (shard.SymmetricGather: peer bulk copies + one barrier, no collective);
    # "nccl": ncclAllGather behind the tower.  Falls back to nccl (and says so) if symmetric memory cannot be set up.
    sg, gather_how = None, "none (1 GPU)"
    if world > 1:
        gather_how = "nccl all-gather behind the tower"
        if args.gather == "symmetric":
            try:
                from video_r2_b200.shard import SymmetricGather
                cap = max(tokens_total, 0 if args.no_side_configs else max(grid_tokens(g_) for g_ in SIDE_CONFIGS.values()))
                sg = SymmetricGather(tower, cap, depth=2)
                gather_how = "fused: peer stores from the merger's last GEMM (symmetric memory) + one barrier"
            except Exception as e:
                print(f"[bench] symmetric memory unavailable ({e!r}); using the NCCL all-gather", file=sys.stderr)
                gather_how += f" (symmetric memory unavailable: {type(e).__name__})"

    def step_device(i):
        if sg is not None:
            full = sg.encode_share(pix_dev[i % n_rot], me, tokens_total)
            local = full[me.token_lo:me.token_hi]
        else:
            local = tower(pix_dev[i % n_rot], my_grid)
            full = all_gather_ragged(local, counts) if world > 1 else local
        pos, delta = vr2.get_rope_index(ids_dev, None, vid_thw, spg, None)
        return local, full, pos

    def step_e2e(i):
        x = pix_host[i % n_rot].to(dev, non_blocking=True)
        ids = ids_host.to(dev, non_blocking=True)
        if sg is not None:
            full = sg.encode_share(x, me, tokens_total)
            local = full[me.token_lo:me.token_hi]
        else:
            local = tower(x, my_grid)
            full = all_gather_ragged(local, counts) if world > 1 else local
        pos, delta = vr2.get_rope_index(ids, None, vid_thw, spg, None)
        out_host.copy_(local, non_blocking=True)
        pos_host.copy_(pos, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller consumes the result
        return full

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(max(3, args.warmup)):
        step_device(i)
    # ---- value: inputs resident in HBM; the dominant kernel is timed live with CUDA events
    tower.profile(["gate_up"])
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop()
    prof = tower.profile_read()
    tower.profile(())
    ms_step = ms_total / args.steps
    value = tokens_total / (ms_step * 1e-3)
    # ---- e2e: host buffers in, host buffers out, through the public host-to-host API (video_r2_b200.HostPipeline):
    #      every step's pixels + input_ids are copied from pinned host memory and its embeddings + position ids are
    #      copied back inside the timed region; copies of neighbouring steps overlap the kernels (3 streams).
    #      e2e_sync is the same work with a synchronous .to(device) -> forward -> .cpu() per step.
    pipe = vr2.HostPipeline(tower, depth=2, gather_counts=counts if world > 1 else None, symmetric=sg,
                            share=me if sg is not None else None)
    out_hosts = [torch.empty((tok_local, cfg.out_hidden_size), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    pos_hosts = [torch.empty((3, world, len(row)), dtype=torch.int64).pin_memory() for _ in range(2)]

    def run_pipe(steps):
        res = pipe.results()
        for i in range(steps):
            pipe.submit(pix_host[i % n_rot], my_grid, out_host=out_hosts[i % 2], input_ids_host=ids_host,
                        video_grid_thw=vid_thw, second_per_grid_ts=spg, pos_host=pos_hosts[i % 2])
            if i >= 1:
                next(res)        # the consumer runs one step behind: step i-1's host buffers are complete here
        pipe.drain()

    def timed_pipe(fn, steps):
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)                # drain() returns when the last step's D2H has landed in host memory
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), wall)   # the copies run on side streams: the host clock bounds them too
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # the same from resized uint8 FRAMES (what decord + the CPU resize produce): 1 byte per pixel value up, rescale +
    # normalise + patchify on the GPU (vr2.patchify, bit-exact with the HF processor) -- reported as e2e.from_frames
    frames_host = [torch.randint(0, 256, (2 * me.grid[0][0] if len(me.grid) == 1 else 2, 3, CLIP[1] * 14, CLIP[2] * 14),
                                 dtype=torch.uint8, generator=torch.Generator().manual_seed(7 + k)).pin_memory()
                   for k in range(2)] if len(me.grid) == 1 else None

    def run_pipe_frames(steps):
        res = pipe.results()
        for i in range(steps):
            pipe.submit_frames(frames_host[i % 2], out_host=out_hosts[i % 2], input_ids_host=ids_host,
                               video_grid_thw=vid_thw, second_per_grid_ts=spg, pos_host=pos_hosts[i % 2])
            if i >= 1:
                next(res)
        pipe.drain()

    # and from the decoder's frames at SOURCE resolution, [T,H,W,3] uint8 as decord hands them over: fetch_video's
    # anti-aliased bicubic resize (vr2.resize_frames, byte-identical to torchvision on the CPU) also moves to the GPU
    RAW = (720, 720)
    raw_host = [torch.randint(0, 256, (frames_host[0].shape[0], RAW[0], RAW[1], 3), dtype=torch.uint8,
                              generator=torch.Generator().manual_seed(17 + k)).pin_memory().permute(0, 3, 1, 2)
                for k in range(2)] if frames_host is not None else None

    def run_pipe_raw(steps):
        res = pipe.results()
        for i in range(steps):
            pipe.submit_frames(raw_host[i % 2], size=(CLIP[1] * 14, CLIP[2] * 14), out_host=out_hosts[i % 2],
                               input_ids_host=ids_host, video_grid_thw=vid_thw, second_per_grid_ts=spg,
                               pos_host=pos_hosts[i % 2])
            if i >= 1:
                next(res)
        pipe.drain()

    run_pipe(3)
    ms_e2e = timed_pipe(run_pipe, args.steps) / args.steps
    e2e_value = tokens_total / (ms_e2e * 1e-3)
    ms_e2e_frames = ms_e2e_raw = None
    if frames_host is not None and world == 1:
        run_pipe_frames(3)
        ms_e2e_frames = timed_pipe(run_pipe_frames, args.steps) / args.steps
        run_pipe_raw(3)
        ms_e2e_raw = timed_pipe(run_pipe_raw, args.steps) / args.steps
    for i in range(2):
        step_e2e(i)
    ms_e2e_sync = timed(step_e2e, args.steps) / args.steps
    h2d = pix_host[0].numel() * 2 + ids_host.numel() * 8
    d2h = out_host.numel() * 2 + pos_host.numel() * 8
    # ---- per-kind breakdown (untimed pass, all kinds) for the profile notes
    tower.profile(tower.KERNEL_KINDS)
    for i in range(3):
        step_device(i)
    breakdown = tower.profile_read()
    tower.profile(())

    # ---- the other operand precisions on the same workload (include/vr2.h vr2_precision): "split" (hi + lo bf16, 2x the
    #      MMA work) and "f16" (fp16 activations) meet max-abs 2e-2 against the fp32 reference (tests/test_gpu_tower.py);
    #      "bf16" is HF's own bf16 arithmetic.  Differences are reported against the split output of the same input.
    precise = None
    modes = None
    if not args.no_precise:
        DESC = {"bf16": "bf16 activations x bf16 weights (HF's bf16 arithmetic)",
                "f16": "fp16 activations x fp16 copies of the bf16 weights (same MMA work as bf16)",
                "auto": "f16 + the saturation guard read (D2H of the counter + stream sync) after every forward, split re-run "
                        "if it tripped: the library's shipped default",
                "split": "hi+lo bf16 activations x bf16 weights (2x MMA work)"}
        outs, modes = {}, {}
        outs[args.precision] = tower(pix_dev[0], my_grid, out_dtype=torch.float32).clone()
        for mode in ("bf16", "f16", "auto", "split"):
            if mode == args.precision:
                continue
            tower.set_precision(mode)
            for i in range(3):
                step_device(i)
            ms_p = timed(step_device, args.steps) / args.steps
            outs[mode] = tower(pix_dev[0], my_grid, out_dtype=torch.float32).clone()
            modes[mode] = {"operands": DESC[mode], "value": tokens_total / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p,
                           "step_tflops_algorithmic": tokens_total * FLOP_PER_TOKEN / (ms_p * 1e-3) / 1e12 / world}
        tower.set_precision(args.precision)
        modes[args.precision] = {"operands": DESC[args.precision], "value": value, "unit": UNIT, "ms_per_step": ms_step,
                                 "headline": True}
        for mode, o in outs.items():
            if mode != "split":
                modes[mode]["max_abs_vs_split"] = (o - outs["split"]).abs().max().item()
                modes[mode]["cosine_vs_split"] = torch.nn.functional.cosine_similarity(o.flatten(), outs["split"].flatten(), dim=0).item()
        precise = dict(modes["split"], precision="split")
        del outs

    # ---- SURVEY §8f rank 1 (the caller side): tower + masked_scatter into inputs_embeds, fused vs the reference's way
    widening = None
    if world == 1:
        emb = torch.zeros((1, len(row), cfg.out_hidden_size), dtype=torch.bfloat16, device=dev)
        vmask = (ids_dev[:1] == 151656).unsqueeze(-1).expand_as(emb)

        def ref_way(i):   # monkey_patch_forward.py:392-406 with the tower already replaced
            feats = tower(pix_dev[i % n_rot], my_grid)
            n_tok = (ids_dev[:1] == 151656).sum().item()
            assert n_tok == feats.shape[0]
            return emb.masked_scatter(vmask, feats)

        def fused_way(i):
            return vr2.embed_visual_tokens(tower, emb, ids_dev[:1], 151656, pix_dev[i % n_rot], my_grid)

        for fn in (ref_way, fused_way):
            for i in range(2):
                fn(i)
        # alternate the two (the step is power-capped: whatever runs later in a long burst runs at a lower clock)
        t_ref, t_fused = [], []
        for _ in range(3):
            t_fused.append(timed(fused_way, args.steps) / args.steps)
            t_ref.append(timed(ref_way, args.steps) / args.steps)
        ms_ref, ms_fused = statistics.median(t_ref), statistics.median(t_fused)
        same = torch.equal(ref_way(0), fused_way(0).clone())
        widening = {"what": "tower + placement of the visual tokens in inputs_embeds [1, %d, %d] bf16" % (len(row), cfg.out_hidden_size),
                    "forward_then_masked_scatter_ms": ms_ref, "fused_scatter_ms": ms_fused, "bit_identical": bool(same)}
        # LM-head per-token log-probs without the logits tensor, vs matmul + log_softmax + gather
        try:
            Mt, Kt, Vt = 4096, 3584, 152064
            gl = torch.Generator(device=dev).manual_seed(9)
            hl = torch.randn(Mt, Kt, generator=gl, device=dev).to(torch.bfloat16)
            wl = (torch.randn(Vt, Kt, generator=gl, device=dev) * 0.05).to(torch.bfloat16)
            tl = torch.randint(0, Vt, (Mt,), generator=gl, device=dev)

            def lp_fused(i):
                return vr2.lm_logprob(hl, wl, tl, 1.0)

            def lp_torch(i):
                logits = hl @ wl.T
                return torch.log_softmax(logits.float(), -1).gather(-1, tl[:, None]).squeeze(-1)

            for fn in (lp_fused, lp_torch):
                fn(0)
            ms_lf = timed(lp_fused, 5) / 5
            ms_lt = timed(lp_torch, 5) / 5
            diff = (lp_fused(0) - lp_torch(0)).abs().max().item()
            widening["lm_logprob"] = {"tokens": Mt, "vocab": Vt, "fused_ms": ms_lf, "matmul_log_softmax_ms": ms_lt,
                                      "fused_tflops": 2.0 * Mt * Kt * Vt / (ms_lf * 1e-3) / 1e12, "max_abs_diff": diff}
            del hl, wl, tl
        except Exception as e:  # never let the optional measurement break the contract line
            widening["lm_logprob"] = {"error": repr(e)}
        # the decoder's q|k|v projection with M-RoPE fused into the GEMM epilogue, vs what the
        # reference runs (three cuBLAS Linears + Hugging Face's rotary functions), Qwen2.5-VL-7B decoder geometry
        try:
            from transformers.models.qwen2_5_vl import modeling_qwen2_5_vl as HFM
            Bq, Sq, Kq, Hq, Hk, Dq = 2, 4096, 3584, 28, 4, 128
            gq = torch.Generator(device=dev).manual_seed(10)
            hid = torch.randn(Bq, Sq, Kq, generator=gq, device=dev).to(torch.bfloat16)
            lins = [torch.nn.Linear(Kq, n, bias=True, device=dev, dtype=torch.bfloat16) for n in (Hq * Dq, Hk * Dq, Hk * Dq)]
            posq = torch.randint(0, 6000, (3, Bq, Sq), generator=gq, device=dev)
            invq = (1.0 / (1000000.0 ** (torch.arange(0, Dq, 2, dtype=torch.int64).float() / Dq))).to(dev)
            fusedq = vr2.FusedDecoderQKV(*lins, head_dim=Dq)

            def qkv_fused(i):
                c, s_ = vr2.mrope_table(posq, invq)
                return fusedq(hid, c, s_)

            def qkv_ref(i):    # Qwen2_5_VLRotaryEmbedding.forward + Qwen2_5_VLAttention.forward up to the rotary
                with torch.no_grad():
                    fr = (invq[None, None, :, None].float().expand(3, Bq, -1, 1) @ posq[:, :, None, :].float()).transpose(2, 3)
                    emb = torch.cat((fr, fr), dim=-1)
                    c, s_ = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
                    q_, k_, v_ = (l(hid).view(Bq, Sq, -1, Dq).transpose(1, 2) for l in lins)
                    q_, k_ = HFM.apply_multimodal_rotary_pos_emb(q_, k_, c, s_, [16, 24, 24])
                    return q_, k_, v_

            for fn in (qkv_fused, qkv_ref):
                fn(0)
            ms_qf = timed(qkv_fused, 5) / 5
            ms_qr = timed(qkv_ref, 5) / 5
            dq = max((a_.float() - b_.float()).abs().max().item() for a_, b_ in zip(qkv_fused(0), qkv_ref(0)))
            widening["decoder_qkv_mrope"] = {"tokens": Bq * Sq, "fused_ms": ms_qf, "linear_x3_plus_hf_rotary_ms": ms_qr,
                                             "fused_tflops": 2.0 * Bq * Sq * Kq * (Hq + 2 * Hk) * Dq / (ms_qf * 1e-3) / 1e12,
                                             "max_abs_diff_vs_bf16_pipeline": dq}
            del hid, lins, fusedq
        except Exception as e:
            widening["decoder_qkv_mrope"] = {"error": repr(e)}

    # ---- the other BASELINE configs (outside the headline's timed region).  N = 1: the whole batch on this GPU.
    #      N > 1: STRONG-scaled through shard.encode_sharded — every rank holds the same pixels (GRPO's RepeatSampler
    #      hands all ranks the same prompt, grpo_trainer.py:664-679), encodes its frame-balanced share of the temporal
    #      slices and one (ragged) all-gather returns the full embeddings to every rank; gather_check: rank 0 encodes the
    #      first slices of the LAST rank's share alone and compares the gathered rows bit for bit.
    side = {}
    if not args.no_side_configs:
        from video_r2_b200.shard import encode_sharded
        names = ["cfg3_rollout_8x64f_448", "cfg5_ragged_16_clips"] if world > 1 else list(SIDE_CONFIGS)
        if world == 1:
            SIDE_CONFIGS["cfg3_one_clip_per_gpu_64f_448"] = [[32, 32, 32]]
            names = ["cfg1_16f_448", "cfg3_one_clip_per_gpu_64f_448", "cfg4_128f_336x588", "cfg5_ragged_16_clips"]
        del pix_dev[1:]
        torch.cuda.empty_cache()
        for name in names:
            g = SIDE_CONFIGS[name]
            try:
                n_p = sum(t * h * w for t, h, w in g)
                gen = torch.Generator(device=dev).manual_seed(1234)      # the same pixels on every rank
                xs = torch.randn(n_p, cfg.patch_dim, generator=gen, device=dev, dtype=torch.bfloat16)
                sh = partition_slices(g, world)
                costs = [s_.cost for s_ in sh]

                def run(i):
                    if sg is not None:
                        return sg.encode(xs, g)
                    return encode_sharded(tower, xs, g) if world > 1 else tower(xs, g)

                k = max(2, min(5, args.steps))
                for i in range(2):
                    full = run(i)
                ms = timed(run, k) / k
                entry = {"grid_thw": g if len(g) <= 2 else f"{len(g)} clips", "patches": n_p, "tokens": n_p // 4,
                         "ms": ms, "tokens_per_sec": n_p / 4 / (ms * 1e-3), "steps": k,
                         "tflops_per_gpu": grid_flops(g) / (ms * 1e-3) / 1e12 / world,
                         "scaling": "strong" if world > 1 else "single GPU"}
                if world > 1:
                    entry["cost_imbalance_max_over_mean"] = max(costs) / (sum(costs) / len(costs))
                    entry["slices_per_rank"] = [sum(r[0] for r in s_.grid) for s_ in sh]
                    entry["tokens_per_rank"] = [s_.token_hi - s_.token_lo for s_ in sh]
                    last = [s_ for s_ in sh if s_.patch_hi > s_.patch_lo][-1]
                    t0, h0, w0 = last.grid[0]
                    sub = [[min(4, t0), h0, w0]]
                    n_sub = sub[0][0] * h0 * w0
                    ok = True
                    if rank == 0:
                        alone = tower(xs[last.patch_lo:last.patch_lo + n_sub], sub)
                        ok = bool(torch.equal(full[last.token_lo:last.token_lo + n_sub // 4], alone))
                    entry["gather_check"] = ok
                    entry["gather_check_what"] = (f"rows {last.token_lo}..{last.token_lo + n_sub // 4} of the gathered tensor (rank "
                                                  f"{sh.index(last)}'s first {sub[0][0]} slices) == the same slices encoded alone on rank 0, torch.equal")
                side[name] = entry
                del xs, full
            except Exception as e:   # an optional line must not cost the contract line
                side[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        sat = tower.saturation_count() if args.precision in ("f16", "auto") else 0
    else:
        sat = tower.saturation_count() if args.precision in ("f16", "auto") else 0

    # ---- the LIBRARY bar on the same GPU: Hugging Face's own module, bf16, flash-attn 2 and sdpa —
    #      what a Video-R2 user runs today (cuBLAS GEMMs + FA2 mma.sync kernels + eager elementwise)
    library_bar = None
    if world == 1 and not args.no_library_bar:
        library_bar = {}
        for impl in ("flash_attention_2", "sdpa"):
            try:
                m = build_hf_reference(attn=impl, device=dev, dtype=torch.bfloat16)
                gthw = torch.tensor([list(CLIP)], device=dev)
                xs = torch.randn(CLIP[0] * CLIP[1] * CLIP[2], cfg.patch_dim, device=dev, dtype=torch.bfloat16)

                def hf(i):
                    with torch.no_grad():
                        return m(xs, grid_thw=gthw)

                for i in range(2):
                    hf(i)
                ms = timed(hf, 3) / 3
                library_bar[impl] = {"ms_per_step": ms, "tokens_per_sec": per_clip / (ms * 1e-3)}
                del m, xs
            except Exception as e:
                library_bar[impl] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
        library_bar["what"] = ("transformers Qwen2_5_VisionTransformerPretrainedModel in bf16 on the same B200, same clip "
                               "([16,32,32]), random init, 3 steps after 2 warm-ups")

    peaks, peak_src = _peaks()
    gu_ms, gu_n = prof.get("gate_up", (0.0, 0))
    gu_avg = gu_ms / max(1, gu_n)
    gu_flops = 2.0 * n_local * (2 * cfg.intermediate_size) * cfg.hidden_size  # algorithmic (unpadded) per launch
    achieved = gu_flops / (gu_avg * 1e-3) / 1e12 if gu_avg > 0 else None
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_top_kernel.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    launches = tower.launch_count(my_grid) + 2
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "f16": "f16", "split": "bf16x2", "auto": "f16"}[args.precision],
        "data": "synthetic",
        "config": {
            "workload": workload_desc(world),
            "frames_per_sec": value / per_clip * 2 * CLIP[0],
            "l2": "pixel inputs rotate over 4 buffers (154 MB > 126 MB L2); 1.35 GB of weights and 0.44 GB of "
                  "activations stream through HBM every step",
            "operands": {"bf16": "bf16 GEMM/attention operands", "f16": "fp16 activations x fp16 copies of the bf16 weights",
                         "auto": "fp16 activations x fp16 copies of the bf16 weights, guard read after every forward",
                         "split": "hi+lo bf16 (split) operands"}[args.precision] + ", fp32 accumulate, fp32 residual stream",
            "precision": args.precision,
            "step_tflops": tokens_total * FLOP_PER_TOKEN / (ms_step * 1e-3) / 1e12 / world,
            "gather": gather_how,
            "host_cores_rank0": (f"{len(host_cores)} cores local to the GPU (NVML affinity)" if host_cores else "not bound"),
        },
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "kernel": "gate|up GEMM + SwiGLU epilogue (42.6% of FLOPs)", "launches_timed": gu_n,
                     "avg_launch_ms": gu_avg, "peak_source": peak_src + ", sustained bf16"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "api": "video_r2_b200.HostPipeline (pinned host in/out, copies overlapped)",
                "sync_value": tokens_total / (ms_e2e_sync * 1e-3), "sync_ms_per_step": ms_e2e_sync,
                "from_frames": None if ms_e2e_frames is None else {
                    "value": tokens_total / (ms_e2e_frames * 1e-3), "ms_per_step": ms_e2e_frames,
                    "h2d_bytes_per_step": int(frames_host[0].numel()) + ids_host.numel() * 8,
                    "what": "uint8 frames [32,3,448,448] up, rescale+normalise+patchify on the GPU, then the same path"},
                "from_raw_frames": None if ms_e2e_raw is None else {
                    "value": tokens_total / (ms_e2e_raw * 1e-3), "ms_per_step": ms_e2e_raw,
                    "h2d_bytes_per_step": int(raw_host[0].numel()) + ids_host.numel() * 8,
                    "what": "decoder frames [32,720,720,3] uint8 up, bicubic-antialias resize to 448x448 + "
                            "rescale+normalise+patchify on the GPU, then the same path"}},
        "gpu_launches": launches * args.steps,
        "clocks": clocks,
        "breakdown_ms_per_step": {k: round(v[0] / 3, 4) for k, v in breakdown.items()},
    }
    if precise is not None:
        line["precise"] = precise
        line["modes"] = modes
    if widening is not None:
        line["widening"] = widening
    if side:
        line["configs"] = side
    if library_bar is not None:
        line["library_bar"] = library_bar
    line["saturation_count"] = sat   # f16 range guard over every forward of this run (0: nothing was clamped)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            kind, run = cpu_reference_runner()
            run(1)
            slices = 2
            dt = run(slices)
            tok = slices * CLIP[1] * CLIP[2] // 4
            line["cpu_baseline"] = {"value": tok / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": f"{slices} of the {CLIP[0]} temporal slices of the same clip ({tok} tokens), "
                                              f"one forward after one warm-up, fp32 sdpa, {dt:.1f} s"}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e!r}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vr2", choices=["vr2", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="f16", choices=["bf16", "f16", "split", "auto"],
                    help="operand precision of the headline numbers (include/vr2.h vr2_precision).  Default f16: the fastest "
                         "mode that meets the north-star bound (max-abs 2e-2 vs the fp32 reference); bf16 (the library "
                         "default, HF's own bf16 arithmetic) and split are measured beside it under `modes`")
    ap.add_argument("--no-precise", action="store_true", help="skip the extra split-precision measurement")
    ap.add_argument("--gather", default="symmetric", choices=["symmetric", "nccl"],
                    help="N > 1: fused peer-store gather from the merger's last GEMM (default) or an NCCL all-gather behind the tower")
    ap.add_argument("--no-side-configs", action="store_true", help="skip the BASELINE configs 1/3/4/5 lines under `configs`")
    ap.add_argument("--no-library-bar", action="store_true", help="skip timing HF bf16 (FA2 / sdpa) on the same GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: relaunch under torchrun
            import subprocess
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        args.gpus = world
    run_vr2(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
