import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: a skipped parity test is not parity.
    pass


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def hf_available():
    try:
        import transformers  # noqa: F401
        from transformers.models.qwen2_5_vl.modeling_qwen2_5_vl import (  # noqa: F401
            Qwen2_5_VisionTransformerPretrainedModel,
        )
        return True
    except Exception:
        return False
