"""Multi-GPU path: shard a batch of clips by TEMPORAL SLICE across the ranks of one node.

A temporal slice (2 frames, h*w patches) is the independent unit of the tower: patch embed,
window attention, the "full" attention layers (one cu_seqlens segment per slice, HF
modeling_qwen2_5_vl.py:488-496), the MLP and the 2x2 merger never cross slices (SURVEY §8e;
tests/test_oracle_tower.py::test_slice_independence).  So the ranks need no exchange while
encoding; the only collective is one all-gather of the merged embeddings at the end, which is
what lets every rank of a GRPO step (RepeatSampler hands all 8 ranks the same prompt,
train/src/trainer/grpo_trainer.py:667-679) encode 1/R of the clip instead of all of it.

The partition is contiguous in the global slice order and balanced on a FLOP model, so each
rank's pixels and each rank's output tokens are ONE contiguous row range: the gathered buffer
lands in original token order without an index pass.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

# FLOPs per patch: dense layers + 28 window-attention layers are constant per patch,
# the 4 full-attention layers cost 4*L*1280 each with L = h*w.
_C_CONST = 1.315e9 - 4 * 4 * 1280 * 1024
_C_FULL = 4 * 4 * 1280


def slice_cost(h: int, w: int) -> float:
    L = h * w
    return L * (_C_CONST + _C_FULL * L)


@dataclasses.dataclass(frozen=True)
class RankShare:
    grid: Tuple[Tuple[int, int, int], ...]  # local grid_thw rows (clips cut at slice boundaries)
    patch_lo: int
    patch_hi: int
    token_lo: int
    token_hi: int
    cost: float


def partition_slices(grid_thw: Sequence[Sequence[int]], world: int, merge: int = 2) -> List[RankShare]:
    """Contiguous, cost-balanced split of all temporal slices over `world` ranks.
    Ranks may receive nothing when there are fewer slices than ranks."""
    slices = []  # (clip, h, w, cost)
    for c, (t, h, w) in enumerate(grid_thw):
        for _ in range(int(t)):
            slices.append((c, int(h), int(w), slice_cost(int(h), int(w))))
    total = sum(s[3] for s in slices)
    # slice i goes to the rank whose cost interval contains the slice's midpoint: monotone, contiguous,
    # balanced to within half a slice
    owner = []
    acc = 0.0
    for _, _, _, k in slices:
        owner.append(min(world - 1, int((acc + k / 2) / total * world)))
        acc += k
    shares: List[RankShare] = []
    i = 0
    patch = token = 0
    for r in range(world):
        rows: List[List[int]] = []
        cost = 0.0
        p0, t0 = patch, token
        while i < len(slices) and owner[i] == r:
            c, h, w, k = slices[i]
            if rows and rows[-1][3] == c:
                rows[-1][0] += 1
            else:
                rows.append([1, h, w, c])
            cost += k
            patch += h * w
            token += (h // merge) * (w // merge)
            i += 1
        shares.append(RankShare(tuple((t, h, w) for t, h, w, _ in rows), p0, patch, t0, token, cost))
    assert i == len(slices)
    return shares


def all_gather_ragged(local: torch.Tensor, counts: Sequence[int], group=None, out: torch.Tensor = None) -> torch.Tensor:
    """One all-gather of per-rank row blocks of different heights.  local: [counts[rank], D].
    Returns [sum(counts), D] in rank order (== original token order for partition_slices); `out` may supply it.

    Every rank's block is one contiguous row range of the result, so nothing is padded or re-copied: equal heights go
    through ncclAllGather, unequal ones through one grouped all-gather-v (torch's NCCL `all_gather` with output views of
    different sizes issues ncclBroadcast per rank inside one ncclGroup), each landing directly in its rows of the
    result.  (Gloo, which the CPU tests of the routing use, has no uneven all-gather: max-padded blocks there.)"""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    D = local.shape[1]
    total = sum(counts)
    if out is None:
        out = local.new_empty((total, D))
    assert out.shape == (total, D) and out.dtype == local.dtype and out.is_contiguous()
    if total == 0:
        return out
    local = local.contiguous()
    if all(c == counts[0] for c in counts):
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    if dist.get_backend(group) == "nccl":
        dist.all_gather([out[offs[r]:offs[r + 1]] for r in range(world)], local, group=group)
        return out
    mx = max(counts)
    padded = local.new_zeros((mx, D))
    padded[: local.shape[0]] = local
    buf = local.new_empty((world, mx, D))
    dist.all_gather_into_tensor(buf.view(world * mx, D), padded, group=group)
    for r in range(world):
        out[offs[r]:offs[r + 1]] = buf[r, : counts[r]]
    return out


class SymmetricGather:
    """The all-gather of the sharded mode WITHOUT a collective: every rank owns `depth` buffers [capacity, D] bf16 in
    symmetric memory (torch.distributed._symmetric_memory: the same allocation on every GPU of the node, each mapped
    into every process over NVLink), and the merger's last GEMM of rank r stores its token rows straight into rows
    [token_lo_r, token_hi_r) of EVERY rank's buffer (`VisionTower.forward_multi`).  What is left of the collective is
    one barrier across the ranks after the tower (signal pads in the same symmetric allocation, enqueued on the
    compute stream; no host synchronisation).

    Buffers rotate (depth >= 2): the tensor returned by `encode` stays valid until `depth - 1` further calls have been
    made, for consumers enqueued on the same stream.  One barrier per call is enough: rank r starts writing buffer
    n % depth in call n only after it has passed the barrier of call n - 1, which every rank reaches — in stream order —
    only after its reads of call n - depth's result (the previous tenant of that buffer) have completed."""

    def __init__(self, tower, capacity_tokens: int, group=None, depth: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        self.tower = tower
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("SymmetricGather: at most 8 ranks (one NVLink node)")
        if depth < 2:
            raise ValueError("SymmetricGather: depth >= 2 (the barrier of call n - 1 protects buffer n % depth)")
        self.D = tower.cfg.out_hidden_size
        self.capacity = int(capacity_tokens)
        self.depth = depth
        self._bufs, self._hdls = [], []
        with torch.cuda.device(tower.device):
            for _ in range(depth):
                t = symm_mem.empty((self.capacity, self.D), dtype=torch.bfloat16, device=tower.device)
                h = symm_mem.rendezvous(t, self.group)
                self._bufs.append(t)
                self._hdls.append(h)
        self._n = 0

    def encode(self, pixels: torch.Tensor, grid_thw) -> torch.Tensor:
        """Every rank holds the same `pixels` [N, 1176] / `grid_thw` (GRPO: RepeatSampler hands every rank the same
        prompt).  Returns the full [N/4, D] embeddings (a view of this rank's symmetric buffer), valid in the current
        stream's order."""
        if isinstance(grid_thw, torch.Tensor):
            grid = grid_thw.detach().to("cpu").reshape(-1, 3).tolist()
        else:
            grid = [list(map(int, r)) for r in grid_thw]
        shares = partition_slices(grid, self.world, self.tower.cfg.spatial_merge_size)
        me = shares[self.rank]
        return self.encode_share(pixels[me.patch_lo:me.patch_hi], me, shares[-1].token_hi)

    def encode_share(self, local_pixels: torch.Tensor, share: RankShare, total_tokens: int, check_guard=None) -> torch.Tensor:
        """This rank's part only: `local_pixels` are the rows of `share` (its clips cut at slice boundaries); the
        result rows [share.token_lo, share.token_hi) are written into every rank's buffer, the rest arrives from the
        peers.  Every rank of the group must make the matching call."""
        if total_tokens > self.capacity:
            raise ValueError(f"SymmetricGather: {total_tokens} tokens exceed the capacity {self.capacity}")
        slot = self._n % self.depth
        self._n += 1
        hdl, buf = self._hdls[slot], self._bufs[slot]
        if share.patch_hi > share.patch_lo:
            row_bytes = self.D * 2
            dests = [int(hdl.buffer_ptrs[r]) + share.token_lo * row_bytes for r in range(self.world)]
            self.tower.forward_multi(local_pixels, [list(r) for r in share.grid], dests, check_guard=check_guard)
        hdl.barrier(channel=0)      # every rank's rows have landed everywhere
        return buf[:total_tokens]


def encode_sharded(tower, pixels: torch.Tensor, grid_thw, group=None, out_dtype=None, gather: bool = True):
    """Every rank holds the same `pixels` [N,1176] / `grid_thw`; rank r encodes its slice range and the
    merged embeddings are all-gathered.  Returns [N/4, out_hidden] on every rank (or the local block
    and the shares when gather=False)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if isinstance(grid_thw, torch.Tensor):
        grid = grid_thw.detach().to("cpu").reshape(-1, 3).tolist()
    else:
        grid = [list(map(int, r)) for r in grid_thw]
    shares = partition_slices(grid, world, tower.cfg.spatial_merge_size)
    me = shares[rank]
    if me.patch_hi > me.patch_lo:
        local = tower(pixels[me.patch_lo:me.patch_hi], [list(r) for r in me.grid], out_dtype=out_dtype)
    else:
        dt = out_dtype or (pixels.dtype if pixels.dtype in (torch.float32, torch.bfloat16) else torch.bfloat16)
        local = torch.empty((0, tower.cfg.out_hidden_size), dtype=dt, device=tower.device)
    if not gather or world == 1:
        return (local, shares) if not gather else local
    return all_gather_ragged(local, [s.token_hi - s.token_lo for s in shares], group)
