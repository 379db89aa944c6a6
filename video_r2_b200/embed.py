"""Placing the tower's output for the decoder: the caller side of `self.visual(...)` in Video-R2
(train/src/train/monkey_patch_forward.py:372-406; the stock HF forward does the same):

    n_video_tokens = (input_ids == video_token_id).sum().item()
    if n_video_tokens != video_embeds.shape[0]: raise ValueError(...)
    mask = (input_ids == video_token_id).unsqueeze(-1).expand_as(inputs_embeds)
    inputs_embeds = inputs_embeds.masked_scatter(mask, video_embeds.to(inputs_embeds.dtype))

`embed_visual_tokens` does this without the extra passes: the positions of the pad token come from one stream
compaction (`vr2_token_positions`), and the merger's last GEMM stores each merged token straight into its row of
`inputs_embeds` (`vr2_tower_forward_scatter`: inverse window permute, dtype and scatter fused into the store).  For
embeddings that already exist (a frozen tower encodes the same clip several times per GRPO step:
grpo_trainer.py:667-679,:783,:826-833) `scatter_rows` places them with one row scatter.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib
from .tower import _dtype_code, _stream_ptr


class _Positions:
    """Result of an asynchronous token_positions: `rows` (int32, -1 past the matches) is valid in stream order;
    `count()` synchronises."""

    def __init__(self, rows, scratch, n):
        self.rows, self._scratch, self._n, self._count = rows, scratch, n, None

    def count(self) -> int:
        if self._count is None:
            c = C.c_int64(0)
            with torch.cuda.device(self.rows.device):
                check(lib().vr2_token_positions_count(self._scratch.data_ptr(), self._n, C.byref(c),
                                                      _stream_ptr(self.rows.device)), "vr2_token_positions_count")
            self._count = c.value
        return self._count


def token_positions_async(input_ids: torch.Tensor, token_id: int, capacity: int = None) -> _Positions:
    if not input_ids.is_cuda:
        raise ValueError("token_positions: input_ids must be a CUDA tensor (no CPU path)")
    ids = input_ids.to(torch.int64).contiguous()
    n = ids.numel()
    cap = n if capacity is None else int(capacity)
    pos = torch.empty(max(cap, 1), dtype=torch.int32, device=ids.device)
    scratch = torch.empty(max(lib().vr2_token_positions_scratch_bytes(n), 4), dtype=torch.uint8, device=ids.device)
    with torch.cuda.device(ids.device):
        check(lib().vr2_token_positions(ids.data_ptr(), n, int(token_id), pos.data_ptr(), cap, scratch.data_ptr(),
                                        None, _stream_ptr(ids.device)), "vr2_token_positions")
    return _Positions(pos[:cap], scratch, n)


def token_positions(input_ids: torch.Tensor, token_id: int, capacity: int = None):
    """-> (positions int32 [min(count, capacity)] on input_ids' device, count).  positions[k] is the flat index (over
    the contiguous [B, S] ids) of the k-th occurrence of token_id — the order masked_scatter consumes rows in."""
    p = token_positions_async(input_ids, token_id, capacity)
    count = p.count()
    return p.rows[: min(count, p.rows.numel())], count


def scatter_rows(src: torch.Tensor, dst: torch.Tensor, dest_rows: torch.Tensor) -> torch.Tensor:
    """dst.view(-1, D)[dest_rows[k]] = src[k] (converted to dst's dtype), in place.  src [n, D]; dst [..., D] contiguous."""
    if src.dim() != 2 or not src.is_cuda or not dst.is_cuda or not dst.is_contiguous():
        raise ValueError("scatter_rows: src must be a 2-D CUDA tensor and dst a contiguous CUDA tensor")
    if dst.shape[-1] != src.shape[1]:
        raise ValueError(f"scatter_rows: row width {src.shape[1]} does not match destination width {dst.shape[-1]}")
    src = src.contiguous()
    rows = dest_rows.to(device=src.device, dtype=torch.int32).contiguous()
    if rows.numel() != src.shape[0]:
        raise ValueError(f"scatter_rows: {src.shape[0]} rows but {rows.numel()} destinations")
    with torch.cuda.device(src.device):
        check(lib().vr2_scatter_rows(src.data_ptr(), _dtype_code(src.dtype), dst.data_ptr(), _dtype_code(dst.dtype),
                                     rows.data_ptr(), src.shape[0], src.shape[1], dst.shape[-1],
                                     _stream_ptr(src.device)), "vr2_scatter_rows")
    return dst


class _PlaceVisual(torch.autograd.Function):
    """`inputs_embeds.masked_scatter(mask, feats)` for a graph-attached `inputs_embeds` (trainable `embed_tokens`,
    train_grpo.sh): out of place like the reference, and the gradient at the visual rows is CUT — masked_scatter
    routes those rows' gradient to `feats` (a frozen tower: dropped), never to the embedding table."""

    @staticmethod
    def forward(ctx, inputs_embeds, tower, pixel_values, plan, rows):
        out = inputs_embeds.detach().clone()
        tower.forward_into(pixel_values, plan, out, rows)
        ctx.save_for_backward(rows)
        return out

    @staticmethod
    def backward(ctx, grad):
        (rows,) = ctx.saved_tensors
        g = grad.contiguous().clone()
        zeros = torch.zeros((rows.numel(), g.shape[-1]), dtype=g.dtype, device=g.device)
        scatter_rows(zeros, g, rows)         # rows of -1 (no such token) are skipped by the kernel
        return g, None, None, None, None


def embed_visual_tokens(tower, inputs_embeds: torch.Tensor, input_ids: torch.Tensor, token_id: int,
                        pixel_values: torch.Tensor, grid_thw, what: str = "Video") -> torch.Tensor:
    """The reference's tower call + token count check + masked_scatter for one modality
    (monkey_patch_forward.py:392-406).  inputs_embeds [B, S, D] contiguous (bf16 / fp32, D == tower out_hidden_size),
    input_ids [B, S].

    The count check runs FIRST (one small stream compaction + the same host wait as the reference's
    `.sum().item()`), so a token / feature mismatch raises the reference's ValueError before anything is written.
    Without a graph (`inputs_embeds` detached: generation, frozen embeddings) the visual rows are stored IN PLACE into
    `inputs_embeds` from the merger's last GEMM and the same tensor is returned.  With grad enabled and
    `inputs_embeds.requires_grad` the result is a new tensor (as `masked_scatter` returns) whose backward passes the
    gradient through at text rows and zeroes it at visual rows."""
    if not inputs_embeds.is_contiguous() or inputs_embeds.shape[-1] != tower.cfg.out_hidden_size:
        raise ValueError("embed_visual_tokens: inputs_embeds must be contiguous [..., out_hidden_size]")
    plan = tower.plan(grid_thw)
    pos = token_positions_async(input_ids, token_id, capacity=plan.num_tokens)
    count = pos.count()
    if count != plan.num_tokens:
        raise ValueError(f"{what} features and {what.lower()} tokens do not match: tokens: {count}, features {plan.num_tokens}")
    if torch.is_grad_enabled() and inputs_embeds.requires_grad:
        return _PlaceVisual.apply(inputs_embeds, tower, pixel_values, plan, pos.rows)
    with torch.no_grad():
        tower.forward_into(pixel_values, plan, inputs_embeds, pos.rows)
    return inputs_embeds
