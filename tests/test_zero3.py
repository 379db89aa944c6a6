"""The boundary under DeepSpeed ZeRO-3 (train/scripts/zero3.json:16-27, selected by train_sft.sh:29 and train_grpo.sh:42).

DeepSpeed is NOT installed in this image (and there is no network), so `deepspeed.zero.GatheredParameters` is
emulated here by a fake module with the real one's contract: inside the context every listed parameter's `.data` is
the full tensor, outside it is the empty placeholder ZeRO-3 leaves behind (`torch.empty(0)`), and parameters carry
`ds_id` / `ds_tensor` / `ds_shape`.  What is under test is the patch's own logic: it detects partitioned parameters,
gathers ALL of the tower's parameters (not only the top module's) and snapshots inside the context."""
import contextlib
import sys
import types

import pytest
import torch

from video_r2_b200 import patch as P


class _FakeGathered(contextlib.AbstractContextManager):
    entered = 0

    def __init__(self, params, modifier_rank=None):
        self.params = list(params)
        assert modifier_rank is None      # read-only gather

    def __enter__(self):
        type(self).entered += 1
        for p in self.params:
            p.data = p.ds_full
        return self

    def __exit__(self, *exc):
        for p in self.params:
            p.data = torch.empty(0, dtype=p.dtype)
        return False


def _partition(module):
    """What zero.Init leaves behind: every parameter an empty placeholder + ds_* attributes."""
    for i, p in enumerate(module.parameters()):
        p.ds_full = p.data.clone()
        p.ds_shape = p.data.shape
        p.ds_id = i
        p.ds_tensor = p.data.flatten()[: max(1, p.numel() // 8)].clone()
        p.data = torch.empty(0, dtype=p.dtype)


@pytest.fixture
def fake_deepspeed(monkeypatch):
    ds = types.ModuleType("deepspeed")
    ds.zero = types.SimpleNamespace(GatheredParameters=_FakeGathered)
    monkeypatch.setitem(sys.modules, "deepspeed", ds)
    _FakeGathered.entered = 0
    return ds


def _toy():
    m = torch.nn.Sequential(torch.nn.Linear(6, 4), torch.nn.Sequential(torch.nn.Linear(4, 3)))
    return m


def test_full_parameters_is_a_noop_for_ordinary_modules(fake_deepspeed):
    m = _toy()
    with P.full_parameters(m):
        sd = {k: v.clone() for k, v in m.state_dict().items()}
    assert _FakeGathered.entered == 0
    assert sd["0.weight"].shape == (4, 6)


def test_partitioned_parameters_are_gathered_for_the_snapshot(fake_deepspeed):
    m = _toy()
    want = {k: v.clone() for k, v in m.state_dict().items()}
    _partition(m)
    assert all(v.numel() == 0 for v in m.state_dict().values())      # what round 1 handed to vr2_tower_set_weight
    with P.full_parameters(m):
        got = {k: v.clone() for k, v in m.state_dict().items()}
    assert _FakeGathered.entered == 1
    assert got.keys() == want.keys()
    for k in want:
        assert torch.equal(got[k], want[k]), k                         # nested sub-modules included
    assert all(v.numel() == 0 for v in m.state_dict().values())      # released again


def test_partitioned_without_deepspeed_fails_loudly(monkeypatch):
    monkeypatch.setitem(sys.modules, "deepspeed", None)              # import deepspeed -> ImportError
    m = _toy()
    _partition(m)
    with pytest.raises(RuntimeError, match="ZeRO-3 partitioned"):
        with P.full_parameters(m):
            pass


@pytest.mark.gpu
def test_patched_forward_on_a_zero3_partitioned_tower(fake_deepspeed):
    """End to end on the GPU: an HF tower whose parameters are all ZeRO-3 placeholders, called through the patch
    with the reference's text-only dummy (monkey_patch_forward.py:362-366) and a real clip; equals the same module
    encoded before partitioning."""
    from conftest import hf_available
    if not hf_available():
        pytest.skip("transformers not importable")
    import video_r2_b200 as vr2
    from oracle import tower as T
    from test_gpu_patch import _hf_visual
    m = _hf_visual(2, (1,))
    m.load_state_dict(T.make_weights(T.TowerConfig(depth=2, fullatt_block_indexes=(1,)), seed=31))
    m = m.cuda().to(torch.bfloat16)
    grid = torch.tensor([[2, 10, 14]], device="cuda")
    x = torch.randn(280, 1176, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda()
    vr2.install()
    try:
        ref = m(x, grid_thw=grid)
        ref = ref.pooler_output if hasattr(ref, "pooler_output") else ref
        vr2.refresh(m)
        for i, p in enumerate(m.parameters()):
            p.ds_full, p.ds_shape, p.ds_id = p.data.clone(), p.data.shape, i
            p.ds_tensor = p.data.flatten()[:1].clone()
            p.data = torch.empty(0, dtype=p.dtype, device=p.device)
        global _FakeGathered
        orig_exit = _FakeGathered.__exit__

        def exit_cuda(self, *exc):
            for p in self.params:
                p.data = torch.empty(0, dtype=p.dtype, device="cuda")
            return False
        _FakeGathered.__exit__ = exit_cuda
        try:
            dummy = m(torch.zeros(784, 1176, device="cuda", dtype=torch.bfloat16), grid_thw=torch.tensor([[1, 28, 28]], device="cuda"))
            dummy = dummy.pooler_output if hasattr(dummy, "pooler_output") else dummy
            out = m(x, grid_thw=grid)
            out = out.pooler_output if hasattr(out, "pooler_output") else out
        finally:
            _FakeGathered.__exit__ = orig_exit
        assert _FakeGathered.entered == 1                               # gathered once, on the first call only
        assert dummy.shape == (196, 3584) and torch.isfinite(dummy.float()).all()
        assert torch.equal(out, ref)
    finally:
        vr2.uninstall()
