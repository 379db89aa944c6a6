"""CPU: host logic of the multi-GPU path — slice partition and the ragged all-gather — with a
world_size-2 (and 3) gloo group.  The tower itself is replaced by a deterministic row function, so
this checks routing (who encodes what, where it lands), not arithmetic."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_r2_b200 import shard

CFG5 = [(16, 448, 448), (32, 448, 448), (64, 448, 448), (128, 336, 588), (8, 280, 504), (32, 280, 504),
        (128, 280, 504), (16, 336, 588), (64, 336, 588), (4, 448, 448), (32, 392, 392), (64, 224, 392),
        (16, 504, 280), (128, 224, 224), (2, 392, 392), (48, 364, 644)]


def grid_of(frames, H, W):
    return [frames // 2, H // 14, W // 14]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_covers_everything_once_and_balances(world):
    grid = [grid_of(*c) for c in CFG5]  # BASELINE config 5: ragged mixed-resolution batch
    shares = shard.partition_slices(grid, world)
    assert len(shares) == world
    n_patches = sum(t * h * w for t, h, w in grid)
    assert shares[0].patch_lo == 0 and shares[-1].patch_hi == n_patches
    for a, b in zip(shares, shares[1:]):
        assert a.patch_hi == b.patch_lo and a.token_hi == b.token_lo
    for s in shares:
        assert sum(t * h * w for t, h, w in s.grid) == s.patch_hi - s.patch_lo
        assert (s.patch_hi - s.patch_lo) == 4 * (s.token_hi - s.token_lo)
    # concatenated local grids reproduce the slice sequence of the global grid
    flat = [(h, w) for t, h, w in grid for _ in range(t)]
    got = [(h, w) for s in shares for t, h, w in s.grid for _ in range(t)]
    assert flat == got
    costs = [s.cost for s in shares]
    assert max(costs) <= 1.08 * (sum(costs) / world)  # frame-balanced within a slice


def test_partition_fewer_slices_than_ranks():
    shares = shard.partition_slices([[3, 8, 8]], 8)
    assert sum(s.token_hi - s.token_lo for s in shares) == 3 * 16
    assert sum(1 for s in shares if s.patch_hi > s.patch_lo) == 3


class FakeTower:
    """Row function standing in for the tower: token i of a call = mean of its 4 patch rows' first 8 features."""

    def __init__(self):
        self.cfg = type("C", (), {"spatial_merge_size": 2, "out_hidden_size": 8})()
        self.device = torch.device("cpu")

    def __call__(self, pixels, grid, out_dtype=None):
        assert pixels.shape[0] == sum(t * h * w for t, h, w in grid)
        return pixels[:, :8].reshape(-1, 4, 8).mean(1)


def _worker(rank, world, port, grid, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = sum(t * h * w for t, h, w in grid)
        pixels = torch.arange(n * 16, dtype=torch.float32).reshape(n, 16)
        out = shard.encode_sharded(FakeTower(), pixels, grid)
        ref = FakeTower()(pixels, grid)
        ret[rank] = bool(torch.equal(out, ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, [[4, 8, 8]]), (2, [[3, 8, 12], [2, 4, 4], [5, 6, 8]]), (3, [[2, 4, 4]])])
def test_encode_sharded_gloo(world, grid):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + world * 7 + len(grid)
    mp.spawn(_worker, args=(world, port, grid, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


_SYM_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VR2_ROOT"])
import video_r2_b200 as vr2
from video_r2_b200.shard import SymmetricGather, partition_slices
from oracle import tower as T
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
cfg_o = T.TowerConfig(depth=2, fullatt_block_indexes=(1,))
w = T.make_weights(cfg_o, seed=81)
tower = vr2.VisionTower(vr2.VisionConfig(depth=2, fullatt_block_indexes=(1,)), w, device=dev, precision="f16")
grid = [[3, 8, 12], [2, 4, 6], [1, 26, 46]]
n = sum(t * h * ww for t, h, ww in grid)
x = torch.randn(n, 1176, generator=torch.Generator().manual_seed(82)).to(torch.bfloat16).to(dev)
sg = SymmetricGather(tower, n // 4 + 100, depth=2)
outs = [sg.encode(x, grid).clone() for _ in range(3)]        # rotates through both buffers
whole = tower(x, grid)                                       # the unsharded forward on this rank
shares = partition_slices(grid, world)
for o in outs:
    assert o.shape == whole.shape
    for s in shares:                                         # each rank's rows == that share encoded alone, bit for bit
        if s.patch_hi > s.patch_lo:
            alone = tower(x[s.patch_lo:s.patch_hi], [list(r) for r in s.grid])
            assert torch.equal(o[s.token_lo:s.token_hi], alone), (rank, s)
    # vs the unsharded forward: a slice's rows fall into different 128-row key blocks there (another summation order), and
    # both results are rounded to bf16 — one bf16 ulp of an output of magnitude 4..8 is 0.031
    assert (o.float() - whole.float()).abs().max().item() < 4e-2
    assert torch.nn.functional.cosine_similarity(o.float().flatten(), whole.float().flatten(), dim=0).item() > 0.99999
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("SYMMETRIC_GATHER_OK world", world)
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_symmetric_gather_on_the_gpus_present(tmp_path):
    """shard.SymmetricGather (peer stores from the merger's last GEMM into every rank's symmetric buffer + one barrier)
    with as many ranks as the box has GPUs (1 on the driver's test box: the symmetric-memory rendezvous, the
    multi-destination store and the barrier all still run; 2/4/8 under `gpurun --gpus N`)."""
    import subprocess
    import sys
    import torch
    n = min(torch.cuda.device_count(), 8)
    script = tmp_path / "sym.py"
    script.write_text(_SYM_SCRIPT)
    env = dict(os.environ, VR2_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "SYMMETRIC_GATHER_OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
