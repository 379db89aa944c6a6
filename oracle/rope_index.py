"""ORACLE (test infrastructure, not product code): CPU restatement of
`Qwen2_5_VLForConditionalGeneration.get_rope_index` with the semantics of
transformers==4.51.1, the version Video-R2 pins (requirements.txt:3) and calls at
train/src/train/monkey_patch_forward.py:415-421 / :583-589.

PINNING.  Video-R2 ships no test or golden vector for this function and transformers 4.51.1 itself is not
installed here (no network), so the restatement is pinned against the two LIVE implementations of the same scheme
that are in the image, through fixtures written by tests/golden/make_golden.py (`rope` section):
  * tests/golden/rope_vllm.json — the installed vLLM 0.22 `get_mrope_input_positions` (called unbound on a stub
    `self`), 23 rows incl. the BASELINE config-2 sequence: equal element for element in the multiply-then-truncate
    mode wherever the fp32 product chain (torch: int64 * float32) and vLLM's float64 product truncate alike — every
    element for seconds-per-grid values that are exact in binary;
  * tests/golden/rope_hf55.json — the installed transformers 5.5.0 `get_rope_index` on rows whose vision items have
    t == 1, where its arithmetic still coincides with 4.51.1 (for t > 1 it diverges: SURVEY §8c); unmasked positions
    equal, padding (0 there, 1 in 4.51.1) and the delta's length convention are the documented differences;
plus the worked example of the HF docstring, hand-derived known answers and the fp32 truncation vectors of
SURVEY §8c (tests/test_oracle_rope_index.py).

What no fixture can settle is whether 4.51.1 already contained
`second_per_grid_t = torch.as_tensor(second_per_grid_t, dtype=range_tensor.dtype, device=range_tensor.device)`,
which truncates seconds-per-grid to int64 BEFORE the multiply.  Both behaviours are implemented (`spg_int_cast`),
each bit-exact on the GPU.  The default is the cast, on this evidence: (1) the reference's GRPO / eval callers hand
`second_per_grid_ts` over as a CUDA tensor (grpo_trainer.py:769 `_prepare_inputs`, lmms_eval/models/qwen2_5_vl.py:
302-305 `inputs.to(device)`) while `range_tensor = torch.arange(llm_grid_t)` lives on the CPU — without that
normalising line `cpu_tensor * cuda_0dim_tensor` raises a device-mismatch error, so the version the reference runs
must contain it; (2) the installed 5.5.0 keeps exactly that truncation (`tokens_per_second * int(next(...))`,
modeling_qwen2_5_vl.py:1110).  vLLM follows the original float multiply, which is the other mode.

Written with torch tensor ops on purpose: the temporal index depends on torch's
type promotion (int64 tensor * float32 scalar -> float32), which numpy would not
reproduce.
"""
from __future__ import annotations

import torch

IMAGE_TOKEN = 151655
VIDEO_TOKEN = 151656
VISION_START = 151652


def get_rope_index(
    input_ids,
    image_grid_thw=None,
    video_grid_thw=None,
    second_per_grid_ts=None,
    attention_mask=None,
    *,
    spatial_merge_size: int = 2,
    tokens_per_second=2,
    image_token_id: int = IMAGE_TOKEN,
    video_token_id: int = VIDEO_TOKEN,
    vision_start_token_id: int = VISION_START,
    spg_int_cast: bool = True,
):
    """Returns (position_ids int64 [3,B,S], mrope_position_deltas int64 [B,1])."""
    mrope_position_deltas = []
    if input_ids is not None and (image_grid_thw is not None or video_grid_thw is not None):
        total_input_ids = input_ids
        if attention_mask is None:
            attention_mask = torch.ones_like(total_input_ids)
        position_ids = torch.ones(3, input_ids.shape[0], input_ids.shape[1], dtype=input_ids.dtype)
        image_index, video_index = 0, 0
        for i in range(total_input_ids.shape[0]):
            ids = total_input_ids[i][attention_mask[i] == 1]
            vision_start_indices = torch.argwhere(ids == vision_start_token_id).squeeze(1)
            vision_tokens = ids[vision_start_indices + 1]
            image_nums = int((vision_tokens == image_token_id).sum())
            video_nums = int((vision_tokens == video_token_id).sum())
            input_tokens = ids.tolist()
            llm_pos_ids_list = []
            st = 0
            remain_images, remain_videos = image_nums, video_nums
            for _ in range(image_nums + video_nums):
                if image_token_id in input_tokens and remain_images > 0:
                    ed_image = input_tokens.index(image_token_id, st)
                else:
                    ed_image = len(input_tokens) + 1
                if video_token_id in input_tokens and remain_videos > 0:
                    ed_video = input_tokens.index(video_token_id, st)
                else:
                    ed_video = len(input_tokens) + 1
                if ed_image < ed_video:
                    t, h, w = image_grid_thw[image_index]
                    second_per_grid_t = 0
                    image_index += 1
                    remain_images -= 1
                    ed = ed_image
                else:
                    t, h, w = video_grid_thw[video_index]
                    if second_per_grid_ts is not None:
                        second_per_grid_t = second_per_grid_ts[video_index]
                    else:
                        second_per_grid_t = 1.0
                    video_index += 1
                    remain_videos -= 1
                    ed = ed_video
                llm_grid_t, llm_grid_h, llm_grid_w = int(t), int(h) // spatial_merge_size, int(w) // spatial_merge_size
                text_len = ed - st
                st_idx = llm_pos_ids_list[-1].max() + 1 if len(llm_pos_ids_list) > 0 else 0
                llm_pos_ids_list.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)

                range_tensor = torch.arange(llm_grid_t).view(-1, 1)
                expanded_range = range_tensor.expand(-1, llm_grid_h * llm_grid_w)
                if spg_int_cast:
                    # "normalize type, send to device": int64 cast truncates toward zero
                    second_per_grid_t = torch.as_tensor(second_per_grid_t, dtype=range_tensor.dtype)
                time_tensor = expanded_range * second_per_grid_t * tokens_per_second
                t_index = time_tensor.long().flatten()
                h_index = torch.arange(llm_grid_h).view(1, -1, 1).expand(llm_grid_t, -1, llm_grid_w).flatten()
                w_index = torch.arange(llm_grid_w).view(1, 1, -1).expand(llm_grid_t, llm_grid_h, -1).flatten()
                llm_pos_ids_list.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
                st = ed + llm_grid_t * llm_grid_h * llm_grid_w
            if st < len(input_tokens):
                st_idx = llm_pos_ids_list[-1].max() + 1 if len(llm_pos_ids_list) > 0 else 0
                text_len = len(input_tokens) - st
                llm_pos_ids_list.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)
            llm_positions = torch.cat(llm_pos_ids_list, dim=1).reshape(3, -1)
            position_ids[..., i, attention_mask[i] == 1] = llm_positions
            # 4.51.1: the PADDED row length, not the number of unmasked tokens
            mrope_position_deltas.append(llm_positions.max() + 1 - len(total_input_ids[i]))
        mrope_position_deltas = torch.tensor([int(d) for d in mrope_position_deltas], dtype=torch.int64).unsqueeze(1)
        return position_ids, mrope_position_deltas
    if attention_mask is not None:
        position_ids = attention_mask.long().cumsum(-1) - 1
        position_ids.masked_fill_(attention_mask == 0, 1)
        position_ids = position_ids.unsqueeze(0).expand(3, -1, -1)
        max_position_ids = position_ids.max(0, keepdim=False)[0].max(-1, keepdim=True)[0]
        mrope_position_deltas = max_position_ids + 1 - attention_mask.shape[-1]
    else:
        position_ids = torch.arange(input_ids.shape[1]).view(1, 1, -1).expand(3, input_ids.shape[0], -1)
        mrope_position_deltas = torch.zeros([input_ids.shape[0], 1], dtype=input_ids.dtype)
    return position_ids.contiguous(), mrope_position_deltas


def vllm_style(input_tokens, image_grid_thw, video_grid_thw, second_per_grid_ts, *, spatial_merge_size=2,
               tokens_per_second=2.0, image_token_id=IMAGE_TOKEN, video_token_id=VIDEO_TOKEN,
               vision_start_token_id=VISION_START):
    """Second opinion for a single unpadded row: the algorithm vLLM uses for the same
    model (independent code base, same published scheme; float64 temporal index).
    Returns (positions int64 [3,L], delta int)."""
    import numpy as np

    ids = list(input_tokens)
    L = len(ids)
    starts = [i for i, v in enumerate(ids) if v == vision_start_token_id and i + 1 < L]
    kinds = [ids[i + 1] for i in starts]
    n_img = sum(1 for k in kinds if k == image_token_id)
    n_vid = sum(1 for k in kinds if k == video_token_id)
    chunks = []
    st = 0
    ii = vi = 0
    ri, rv = n_img, n_vid
    for _ in range(n_img + n_vid):
        ed_i = ids.index(image_token_id, st) if (image_token_id in ids and ri > 0) else L + 1
        ed_v = ids.index(video_token_id, st) if (video_token_id in ids and rv > 0) else L + 1
        if ed_i < ed_v:
            t, h, w = [int(v) for v in image_grid_thw[ii]]
            spg = 0.0
            ii += 1
            ri -= 1
            ed = ed_i
        else:
            t, h, w = [int(v) for v in video_grid_thw[vi]]
            spg = float(second_per_grid_ts[vi]) if second_per_grid_ts is not None else 1.0
            vi += 1
            rv -= 1
            ed = ed_v
        gt, gh, gw = t, h // spatial_merge_size, w // spatial_merge_size
        text_len = ed - st
        st_idx = int(chunks[-1].max()) + 1 if chunks else 0
        chunks.append(np.broadcast_to(np.arange(text_len), (3, text_len)) + st_idx)
        t_idx = (np.repeat(np.arange(gt), gh * gw) * spg * tokens_per_second).astype(np.int64)
        h_idx = np.tile(np.repeat(np.arange(gh), gw), gt)
        w_idx = np.tile(np.arange(gw), gt * gh)
        chunks.append(np.stack([t_idx, h_idx, w_idx]) + text_len + st_idx)
        st = ed + gt * gh * gw
    if st < L:
        st_idx = int(chunks[-1].max()) + 1 if chunks else 0
        chunks.append(np.broadcast_to(np.arange(L - st), (3, L - st)) + st_idx)
    pos = np.concatenate(chunks, axis=1).astype(np.int64)
    return pos, int(pos.max()) + 1 - L
