"""ORACLE (test infrastructure, not product code): CPU restatement of how the reference places visual features
into the decoder's input embeddings (Video-R2 train/src/train/monkey_patch_forward.py:372-406; the stock HF forward
modeling_qwen2_5_vl.py has the same block):

    n_tokens = (input_ids == token_id).sum().item()
    if n_tokens != feats.shape[0]: raise ValueError("... features and ... tokens do not match ...")
    mask = (input_ids == token_id).unsqueeze(-1).expand_as(inputs_embeds)
    inputs_embeds = inputs_embeds.masked_scatter(mask, feats.to(inputs_embeds.dtype))

torch.Tensor.masked_scatter copies the elements of `source` in order into the True positions of the (row-major)
flattened mask; with a mask that is constant along the last axis this places source row k at the k-th position whose
id matches.  Pinned against torch.masked_scatter itself in tests/test_oracle_scatter.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module."""
import numpy as np


def token_positions(input_ids, token_id):
    """Flat (row-major) indices of the elements equal to token_id, ascending: masked_scatter's consumption order."""
    ids = np.asarray(input_ids).reshape(-1)
    return np.nonzero(ids == token_id)[0].astype(np.int64)


def embed_visual_tokens(inputs_embeds, input_ids, token_id, feats, what="Video"):
    """numpy arrays: inputs_embeds [B,S,D], input_ids [B,S], feats [n,D] -> new [B,S,D] array (feats cast to its dtype)."""
    pos = token_positions(input_ids, token_id)
    if pos.shape[0] != feats.shape[0]:
        raise ValueError(f"{what} features and {what.lower()} tokens do not match: tokens: {pos.shape[0]}, features {feats.shape[0]}")
    out = np.array(inputs_embeds, copy=True)
    flat = out.reshape(-1, out.shape[-1])
    flat[pos] = feats.astype(out.dtype)
    return out
