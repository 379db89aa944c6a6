"""Host-to-host streaming of the hot path: pinned host pixels (+ input_ids) in, pinned host embeddings
(+ M-RoPE position ids) out, with the copies of neighbouring batches overlapped with the tower's kernels.

Video-R2's callers hand `self.visual(...)` device tensors (the collator has already moved them), so the
drop-in patch never copies.  An evaluation / rollout loop that feeds clips from host memory can use this
instead of a synchronous `.to(device)` -> forward -> `.cpu()` per batch: three CUDA streams (H2D, compute, D2H)
and `depth` device buffers per direction, ordered by events only — no host synchronisation inside `submit`.

    pipe = vr2.HostPipeline(tower, depth=2)
    for x_host, grid in batches:                 # x_host: pinned [N, 1176] bf16 / fp32
        pipe.submit(x_host, grid, out_host=buf)  # returns immediately
    for emb_host, pos_host in pipe.results():    # in submission order; blocks on that batch's D2H event only
        ...
"""
from __future__ import annotations

import collections

import torch

from .preprocess import patchify, resize_frames
from .rope_index import get_rope_index


def bind_to_gpu_node(device) -> list:
    """Restricts this process to the CPU cores NVML reports as local to `device` (its NUMA node), so that the pinned
    host buffers allocated afterwards and the threads that feed the copy engines sit next to the GPU's PCIe root.  With
    one process per GPU on a two-socket host this keeps eight ranks' H2D / D2H traffic off the inter-socket link.
    Returns the core list it bound to ([] when NVML, the affinity call or a usable mask is unavailable: nothing is
    changed then).  Call it once per rank before creating a HostPipeline."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        dev = torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:   # NVML counts physical devices
            idx = int(vis.split(",")[idx])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cores) & os.sched_getaffinity(0))
        if not allowed:
            return []
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return []


class HostPipeline:
    _SNAPS = 256   # guard snapshots kept (= batches that may be in flight)

    def __init__(self, tower, depth: int = 2, rope_config=None, gather_counts=None, group=None, symmetric=None,
                 share=None):
        """gather_counts (tokens per rank, from shard.partition_slices) turns on the SHARDED mode: every batch's
        embeddings are all-gathered over `group` right behind the tower (on the compute stream, overlapping the
        neighbouring batches' copies); `gathered(k)` returns the full [sum(counts), D] device tensor of the batch that
        was submitted k calls ago, and the host result stays this rank's own block."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.gather_counts = None if gather_counts is None else [int(c) for c in gather_counts]
        self.group = group
        # symmetric: a shard.SymmetricGather + this rank's RankShare -> the gather is fused into the tower's last GEMM
        # (peer stores + one barrier) instead of an NCCL all-gather behind it
        self.symmetric, self.share = symmetric, share
        if symmetric is not None and (share is None or gather_counts is None or symmetric.depth < depth):
            raise ValueError("HostPipeline(symmetric=...) needs share=, gather_counts= and symmetric.depth >= depth")
        self._full = [None] * depth        # device buffers of the gathered embeddings (sharded mode)
        self.tower = tower
        self.dev = tower.device
        self.depth = depth
        self.rope_config = rope_config
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._in = [None] * depth          # device pixel / frame buffers
        self._pix = [None] * depth         # device patch rows produced by patchify (submit_frames)
        self._rsz = [None] * depth         # device frames after the resize (submit_frames(size=...))
        self._ids = [None] * depth
        self._out = [None] * depth         # device embedding buffers
        self._compute_done = [None] * depth
        self._d2h_done = [None] * depth
        self._pending = collections.deque()
        self._n = 0
        # f16 range guard (tower precision "auto" / "f16"): the counter is snapshotted behind every forward into pinned
        # host memory; results() compares neighbouring snapshots and repeats a clamped batch in the split precision
        self._guard = tower.precision in ("auto", "f16") and gather_counts is None
        self._snap = torch.zeros((self._SNAPS, 4), dtype=torch.int32).pin_memory() if self._guard else None   # ring, one row per submit
        self._snap_seen, self._seen_epoch, self._epoch = (0, 0, 0, 0), 0, 0   # counters restart (new epoch) after a redo
        if self._guard:
            tower.saturation_count(reset=True)
        self.fallbacks = 0

    def _buf(self, store, slot, shape, dtype):
        b = store[slot]
        if b is None or b.shape != tuple(shape) or b.dtype != dtype:
            b = torch.empty(tuple(shape), dtype=dtype, device=self.dev)
            store[slot] = b
        return b

    @torch.no_grad()
    def submit_frames(self, frames_host: torch.Tensor, size=None, **kw):
        """Like submit(), from uint8 (or, already resized, fp32) frames [T, 3, H, W] in pinned host memory: the upload is
        1 byte per pixel value; the resize (`size=(h', w')`: fetch_video's anti-aliased bicubic, `vr2.resize_frames`) and
        rescale + normalise + patchify (the HF video processor's CPU work, `vr2.patchify`) run on the GPU.  A permuted
        [T, H, W, 3] decoder buffer is accepted as is.  One clip per call; grid_thw is derived from the frame shape."""
        T, _, H, W = frames_host.shape
        if size is not None:
            H, W = int(size[0]), int(size[1])
        grid = [[-(-T // self.tower.cfg.temporal_patch_size), H // self.tower.cfg.patch_size, W // self.tower.cfg.patch_size]]
        return self.submit(frames_host, grid, _frames=True, _size=size, **kw)

    @torch.no_grad()
    def submit(self, x_host: torch.Tensor, grid_thw, out_host: torch.Tensor = None, input_ids_host: torch.Tensor = None,
               video_grid_thw=None, second_per_grid_ts=None, pos_host: torch.Tensor = None, out_dtype=None,
               _frames: bool = False, _size=None):
        """Enqueue one batch.  `grid_thw` describes x_host's rows; when `input_ids_host` is given the M-RoPE position
        ids of that batch are computed too (`video_grid_thw` / `second_per_grid_ts`: device or host tensors)."""
        slot = self._n % self.depth
        self._n += 1
        cur = torch.cuda.current_stream(self.dev)
        plan = self.tower.plan(grid_thw)
        odt = out_dtype or (x_host.dtype if x_host.dtype in (torch.float32, torch.bfloat16) else torch.bfloat16)
        # ---- H2D on the copy-in stream, after the compute that last read this slot
        with torch.cuda.stream(self.s_in):
            if self._compute_done[slot] is not None:
                self.s_in.wait_event(self._compute_done[slot])
            if x_host.dim() == 4 and not x_host.is_contiguous() and x_host.permute(0, 2, 3, 1).is_contiguous():
                # decoder layout [T, H, W, C] viewed as [T, C, H, W]: upload the bytes as they lie, keep the view
                hwc = x_host.permute(0, 2, 3, 1)
                x_dev = self._buf(self._in, slot, hwc.shape, hwc.dtype)
                x_dev.copy_(hwc, non_blocking=True)
                x_dev = x_dev.permute(0, 3, 1, 2)
            else:
                x_dev = self._buf(self._in, slot, x_host.shape, x_host.dtype)
                x_dev.copy_(x_host, non_blocking=True)
            ids_dev = None
            if input_ids_host is not None:
                ids_dev = self._buf(self._ids, slot, input_ids_host.shape, input_ids_host.dtype)
                ids_dev.copy_(input_ids_host, non_blocking=True)
            h2d = torch.cuda.Event()
            h2d.record(self.s_in)
        # ---- kernels on the caller's stream, after the H2D and after the D2H that last read this slot's output
        cur.wait_event(h2d)
        if self._d2h_done[slot] is not None:
            cur.wait_event(self._d2h_done[slot])
        out_dev = self._buf(self._out, slot, (plan.num_tokens, self.tower.cfg.out_hidden_size), odt)
        if _frames:
            cfg = self.tower.cfg
            pix = self._buf(self._pix, slot, (plan.num_patches, cfg.patch_dim), torch.bfloat16)
            if _size is not None:
                rs = self._buf(self._rsz, slot, (x_dev.shape[0], x_dev.shape[1], int(_size[0]), int(_size[1])), torch.uint8)
                x_dev = resize_frames(x_dev, _size, out=rs)
            elif not x_dev.is_contiguous():
                x_dev = x_dev.contiguous()
            patchify(x_dev, patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size,
                     merge_size=cfg.spatial_merge_size, out=pix)
            x_dev = pix
        # sharded mode: a clamped block must be repeated BEFORE it is all-gathered (a rank cannot redo the collective
        # on its own later), so precision "auto" reads the guard synchronously here; otherwise it is deferred to results()
        sync_guard = self.gather_counts is not None and self.tower.precision == "auto"
        if self.symmetric is not None:
            full = self.symmetric.encode_share(x_dev, self.share, sum(self.gather_counts), check_guard=sync_guard)
            self._full[slot] = full
            out_dev = full[self.share.token_lo:self.share.token_hi]
        else:
            self.tower.forward(x_dev, plan, out=out_dev, check_guard=sync_guard)
        if self.gather_counts is not None and self.symmetric is None:
            from .shard import all_gather_ragged
            full = self._buf(self._full, slot, (sum(self.gather_counts), self.tower.cfg.out_hidden_size), odt)
            all_gather_ragged(out_dev, self.gather_counts, self.group, out=full)
        if self._guard:
            if len(self._pending) >= self._SNAPS:
                raise RuntimeError(f"HostPipeline: more than {self._SNAPS} batches in flight; consume results()")
            self.tower.saturation_snapshot(self._snap[(self._n - 1) % self._SNAPS:])
        pos_dev = None
        if ids_dev is not None:
            kw = {"config": self.rope_config} if self.rope_config is not None else {}
            pos_dev, _ = get_rope_index(ids_dev, None, video_grid_thw, second_per_grid_ts, None, **kw)
        done = torch.cuda.Event()
        done.record(cur)
        self._compute_done[slot] = done
        # ---- D2H on the copy-out stream
        if out_host is None:
            out_host = torch.empty(out_dev.shape, dtype=out_dev.dtype).pin_memory()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            out_host.copy_(out_dev, non_blocking=True)
            if pos_dev is not None:
                if pos_host is None:
                    pos_host = torch.empty(pos_dev.shape, dtype=pos_dev.dtype).pin_memory()
                pos_host.copy_(pos_dev, non_blocking=True)
                pos_dev.record_stream(self.s_out)
            d2h = torch.cuda.Event()
            d2h.record(self.s_out)
        self._d2h_done[slot] = d2h
        self._pending.append((d2h, out_host, pos_host, (self._n - 1) % self._SNAPS, self._epoch, (x_host, plan, _frames, _size)))

    @torch.no_grad()
    def _redo_split(self, out_host, x_host, plan, frames, size):
        """Rare path: the f16 guard tripped on this batch.  Recompute it from the host input, synchronously, under the
        tower's "auto" policy (raise the range shifts, else hi+lo bf16 operands), and overwrite its host result."""
        self.fallbacks += 1
        cfg = self.tower.cfg
        x = x_host.to(self.dev)
        if frames:
            if size is not None:
                x = resize_frames(x, size)
            x = patchify(x.contiguous(), patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size,
                         merge_size=cfg.spatial_merge_size)
        out = torch.empty((plan.num_tokens, cfg.out_hidden_size), dtype=out_host.dtype, device=self.dev)
        self.tower.saturation_counts(reset=True)
        self.tower._launch(plan, x.contiguous(), out, None, check_guard=True)
        out_host.copy_(out)
        torch.cuda.current_stream(self.dev).synchronize()
        # the counters restart from zero for everything submitted from now on (a new epoch); batches already in flight
        # keep comparing against their own epoch's snapshots, and are redone in turn if they clamped too
        self._epoch += 1

    def gathered(self, ago: int = 0) -> torch.Tensor:
        """Sharded mode: the all-gathered embeddings (device, valid in the compute stream's order) of the batch
        submitted `ago` submits before the latest one (ago < depth)."""
        if self.gather_counts is None or not 0 <= ago < self.depth or self._n - 1 - ago < 0:
            raise ValueError("gathered(): sharded mode with 0 <= ago < depth and that many batches submitted")
        return self._full[(self._n - 1 - ago) % self.depth]

    def results(self):
        """Yields (embeddings_host, position_ids_host or None) in submission order."""
        while self._pending:
            ev, out_host, pos_host, snap_i, epoch, redo = self._pending.popleft()
            ev.synchronize()
            if self._guard:
                seen = tuple(self._snap[snap_i].tolist())   # cumulative counters right behind this batch's forward
                prev = self._snap_seen if epoch == self._seen_epoch else (0, 0, 0, 0)
                self._snap_seen, self._seen_epoch = seen, epoch
                if seen != prev:
                    self._redo_split(out_host, *redo)
            yield out_host, pos_host

    def drain(self):
        return list(self.results())
