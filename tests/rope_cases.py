"""Shared input builders for the get_rope_index tests (CPU oracle tests and GPU parity tests)."""
import numpy as np
import torch

IMG, VID, VS, VE, PAD = 151655, 151656, 151652, 151653, 151643


def vision_run(kind, grid, merge=2):
    t, h, w = grid
    return [VS] + [kind] * (t * (h // merge) * (w // merge)) + [VE]


def build_row(items, rng, lead=3, gap=2, tail=4):
    """items: list of (kind, grid). Text ids are random ordinary tokens."""
    ids = rng.integers(1000, 5000, size=lead).tolist()
    for kind, grid in items:
        ids += vision_run(kind, grid) + rng.integers(1000, 5000, size=gap).tolist()
    ids += rng.integers(1000, 5000, size=tail).tolist()
    return ids


def pad_batch(rows, side="right"):
    S = max(len(r) for r in rows)
    ids = torch.full((len(rows), S), PAD, dtype=torch.int64)
    mask = torch.zeros((len(rows), S), dtype=torch.int64)
    for i, r in enumerate(rows):
        if side == "right":
            ids[i, : len(r)] = torch.tensor(r)
            mask[i, : len(r)] = 1
        else:
            ids[i, S - len(r):] = torch.tensor(r)
            mask[i, S - len(r):] = 1
    return ids, mask


def case_suite(seed=0):
    """-> list of (name, kwargs for get_rope_index). Covers SURVEY §8c's list: left/right padding,
    mixed image+video rows, text-only rows inside a multimodal batch, two videos in one row,
    fractional second_per_grid_ts, no mask, list-of-floats spg (SFT collator)."""
    rng = np.random.default_rng(seed)
    cases = []
    v1, v2, im1, im2 = [3, 4, 4], [5, 6, 8], [1, 6, 4], [1, 2, 2]
    # single video, no mask, tensor spg
    cases.append(("video_nomask", dict(input_ids=torch.tensor([build_row([(VID, v1)], rng)]), video_grid_thw=torch.tensor([v1]),
                                       second_per_grid_ts=torch.tensor([1.0]))))
    # fractional spg, float32 tensor
    cases.append(("video_frac_spg", dict(input_ids=torch.tensor([build_row([(VID, [12, 4, 4])], rng)]),
                                         video_grid_thw=torch.tensor([[12, 4, 4]]), second_per_grid_ts=torch.tensor([0.7]))))
    cases.append(("video_frac_spg_big", dict(input_ids=torch.tensor([build_row([(VID, [12, 4, 4])], rng)]),
                                             video_grid_thw=torch.tensor([[12, 4, 4]]), second_per_grid_ts=torch.tensor([9.375]))))
    # list-of-floats spg (SFT)
    cases.append(("video_list_spg", dict(input_ids=torch.tensor([build_row([(VID, v2)], rng)]), video_grid_thw=torch.tensor([v2]),
                                         second_per_grid_ts=[2.5])))
    # spg None -> 1.0
    cases.append(("video_spg_none", dict(input_ids=torch.tensor([build_row([(VID, v2)], rng)]), video_grid_thw=torch.tensor([v2]))))
    # image only
    cases.append(("image_only", dict(input_ids=torch.tensor([build_row([(IMG, im1)], rng)]), image_grid_thw=torch.tensor([im1]))))
    # mixed image + video + image in one row, two videos in another, text-only third row; right padded
    rows = [build_row([(IMG, im1), (VID, v1), (IMG, im2)], rng), build_row([(VID, v2), (VID, v1)], rng, lead=0, gap=0, tail=0),
            rng.integers(1000, 5000, size=9).tolist()]
    for side in ("right", "left"):
        ids, mask = pad_batch(rows, side)
        cases.append((f"mixed_batch_{side}pad", dict(input_ids=ids, attention_mask=mask, image_grid_thw=torch.tensor([im1, im2]),
                                                      video_grid_thw=torch.tensor([v1, v2, v1]),
                                                      second_per_grid_ts=torch.tensor([0.5, 1.3333, 2.0]))))
    # bool / int32 masks
    ids, mask = pad_batch(rows, "left")
    cases.append(("mixed_batch_boolmask", dict(input_ids=ids, attention_mask=mask.bool(), image_grid_thw=torch.tensor([im1, im2]),
                                               video_grid_thw=torch.tensor([v1, v2, v1]), second_per_grid_ts=torch.tensor([1.0, 1.0, 3.7]))))
    # GRPO-like: the same prompt repeated, left padded (all rows equal length -> no padding) with a long video
    g = [16, 8, 8]
    row = build_row([(VID, g)], rng, lead=14, gap=0, tail=20)
    ids, mask = pad_batch([row, row], "left")
    cases.append(("grpo_repeat", dict(input_ids=ids, attention_mask=mask, video_grid_thw=torch.tensor([g, g]),
                                      second_per_grid_ts=torch.tensor([0.2133, 0.2133]))))
    # text-only branches (no grids)
    ids, mask = pad_batch([rng.integers(1000, 5000, size=7).tolist(), rng.integers(1000, 5000, size=4).tolist()], "left")
    cases.append(("text_only_mask", dict(input_ids=ids, attention_mask=mask)))
    cases.append(("text_only_nomask", dict(input_ids=ids)))
    return cases
