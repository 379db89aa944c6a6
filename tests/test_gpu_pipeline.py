"""GPU: the host-to-host streaming helper returns exactly what the synchronous path returns."""
import pytest
import torch

from oracle import tower as T

pytestmark = pytest.mark.gpu


def test_host_pipeline_matches_synchronous_path():
    import video_r2_b200 as vr2
    cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), T.make_weights(cfg_o, seed=2), device="cuda:0")
    grids = [[[2, 8, 12]], [[1, 4, 6], [2, 6, 6]], [[2, 8, 12]], [[3, 10, 8]], [[2, 8, 12]]]
    xs, ids = [], []
    for k, g in enumerate(grids):
        n = sum(t * h * w for t, h, w in g)
        xs.append(torch.randn(n, 1176, generator=torch.Generator().manual_seed(k)).to(torch.bfloat16).pin_memory())
        row = [5, 6, 151652] + [151656] * (g[0][0] * g[0][1] * g[0][2] // 4) + [151653, 7]
        ids.append(torch.tensor([row], dtype=torch.int64).pin_memory())
    pipe = vr2.HostPipeline(tower, depth=2)
    for x, g, i in zip(xs, grids, ids):
        pipe.submit(x, g, input_ids_host=i, video_grid_thw=torch.tensor(g[:1]), second_per_grid_ts=torch.tensor([1.5]))
    got = pipe.drain()
    assert len(got) == len(grids)
    for (emb, pos), x, g, i in zip(got, xs, grids, ids):
        ref = tower(x.cuda(), g)
        rpos, _ = vr2.get_rope_index(i.cuda(), None, torch.tensor(g[:1]).cuda(), torch.tensor([1.5]).cuda(), None)
        assert emb.dtype == torch.bfloat16 and torch.equal(emb, ref.cpu())
        assert torch.equal(pos, rpos.cpu())


def test_host_pipeline_from_uint8_frames_matches_processor_then_tower():
    import video_r2_b200 as vr2
    from oracle import patchify as P
    cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), T.make_weights(cfg_o, seed=2), device="cuda:0")
    pipe = vr2.HostPipeline(tower, depth=2)
    clips = [torch.randint(0, 256, (t, 3, h, w), dtype=torch.uint8, generator=torch.Generator().manual_seed(k)).pin_memory()
             for k, (t, h, w) in enumerate([(4, 112, 168), (5, 56, 56), (4, 112, 168)])]
    for c in clips:
        pipe.submit_frames(c)
    got = pipe.drain()
    for (emb, _), c in zip(got, clips):
        rows, grid = P.patchify(c.numpy(), vr2.preprocess.OPENAI_CLIP_MEAN, vr2.preprocess.OPENAI_CLIP_STD)
        ref = tower(torch.from_numpy(rows).to(torch.bfloat16).cuda(), [grid])     # the reference's path: processor, cast, tower
        assert torch.equal(emb, ref.cpu())


def test_host_pipeline_from_decoder_frames_resizes_on_the_gpu():
    """[T,H,W,3] decoder buffers at source resolution: upload as they lie, fetch_video's resize + the processor on the GPU."""
    import video_r2_b200 as vr2
    from oracle import patchify as P
    from oracle import resize as R
    cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
    tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), T.make_weights(cfg_o, seed=2), device="cuda:0")
    pipe = vr2.HostPipeline(tower, depth=2)
    shapes = [(4, 90, 160), (2, 200, 120), (4, 90, 160)]
    clips = [torch.randint(0, 256, (t, h, w, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(k)).pin_memory()
             .permute(0, 3, 1, 2) for k, (t, h, w) in enumerate(shapes)]
    sizes = [vr2.smart_resize(h, w, 28, 4 * 28 * 28, 12 * 28 * 28) for _, h, w in shapes]
    for c, sz in zip(clips, sizes):
        assert not c.is_contiguous()
        pipe.submit_frames(c, size=sz)
    got = pipe.drain()
    for (emb, _), c, sz in zip(got, clips, sizes):
        frames = R.resize_frames(c.contiguous().numpy(), *sz)
        rows, grid = P.patchify(frames, vr2.preprocess.OPENAI_CLIP_MEAN, vr2.preprocess.OPENAI_CLIP_STD)
        ref = tower(torch.from_numpy(rows).to(torch.bfloat16).cuda(), [grid])
        assert emb.shape == ref.shape and torch.equal(emb, ref.cpu())
