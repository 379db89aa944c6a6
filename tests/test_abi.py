"""CPU: the C-ABI library loads and exports every symbol include/vr2.h declares; the ctypes
table in video_r2_b200/_lib.py covers the header one to one.  No compute is called."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "vr2.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vr2_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for must in ("vr2_tower_forward", "vr2_rope_index", "vr2_plan_create", "vr2_gather_rows", "vr2_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from video_r2_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/vr2.h but not exported by libvr2.so"
    assert handle.vr2_abi_version() == 1


def test_ctypes_table_matches_header():
    from video_r2_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback_in_product_path():
    """The package must not import the oracle or torch math fallbacks."""
    pkg = os.path.join(ROOT, "video_r2_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# ", ""), fn


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import video_r2_b200 as vr2
    with pytest.raises(Exception):
        vr2.VisionTower(vr2.VisionConfig(depth=1), {}, device="cuda:0")
    with pytest.raises(Exception):
        vr2.get_rope_index(torch.zeros(1, 4, dtype=torch.int64))


def test_bind_to_gpu_node_changes_nothing_without_a_gpu():
    """HostPipeline's NUMA helper is best effort: without NVML / a visible GPU it returns [] and leaves the affinity."""
    import os
    import video_r2_b200 as vr2
    before = os.sched_getaffinity(0)
    cores = vr2.bind_to_gpu_node("cuda:0")
    assert isinstance(cores, list)
    if not cores:
        assert os.sched_getaffinity(0) == before
    else:   # a GPU box: bound to a subset of what was allowed
        assert set(cores) <= before and os.sched_getaffinity(0) == set(cores)
        os.sched_setaffinity(0, before)
