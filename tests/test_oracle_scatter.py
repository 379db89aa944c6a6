"""CPU: the scatter oracle against torch.masked_scatter itself (the call the reference makes,
monkey_patch_forward.py:385-406), including ragged batches, leading/trailing matches and the mismatch error."""
import numpy as np
import pytest
import torch

from oracle import scatter as S

VID = 151656


@pytest.mark.parametrize("B,L,n_seed", [(1, 40, 0), (3, 57, 1), (4, 128, 2), (2, 1, 3)])
def test_matches_torch_masked_scatter(B, L, n_seed):
    g = torch.Generator().manual_seed(n_seed)
    ids = torch.randint(0, 1000, (B, L), generator=g)
    ids[torch.rand(B, L, generator=g) < 0.4] = VID
    ids[0, 0] = VID
    ids[-1, -1] = VID
    n = int((ids == VID).sum())
    emb = torch.randn(B, L, 16, generator=g)
    feats = torch.randn(n, 16, generator=g)
    mask = (ids == VID).unsqueeze(-1).expand_as(emb)
    ref = emb.masked_scatter(mask, feats)
    out = S.embed_visual_tokens(emb.numpy(), ids.numpy(), VID, feats.numpy())
    assert np.array_equal(out, ref.numpy())
    assert np.array_equal(S.token_positions(ids.numpy(), VID), torch.nonzero((ids == VID).flatten()).flatten().numpy())


def test_mismatch_raises_like_the_reference():
    ids = np.array([[1, VID, VID, 2]])
    with pytest.raises(ValueError, match="features and video tokens do not match"):
        S.embed_visual_tokens(np.zeros((1, 4, 8), np.float32), ids, VID, np.zeros((3, 8), np.float32))
