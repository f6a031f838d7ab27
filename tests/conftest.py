import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _restore_global_modes(request):
    """GPU tests switch library-wide modes (operand precision, lanes, window pairing): put the defaults back afterwards."""
    yield
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            import mumpy_b200
            from mumpy_b200 import ops, streams
            from mumpy_b200.models.encoder import multiTemporalViewEncoder as mtv
            mumpy_b200.set_precision(ops.DEFAULT_PRECISION)
            streams.set_enabled(True)
            mtv.set_per_clip_pairing(False)
            ops.f16_overflowed()
