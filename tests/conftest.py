import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI shared library; built in-tree here when nvcc is around and it is stale."""
    from gapro_b200 import _lib, build
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        build.build()
    return _lib.load()


def oracle_args(inp):
    import numpy as np
    return (inp["xyz"], inp["mask_feats"].astype(np.float32), inp["spp"], inp["instance_cls"].astype(np.int64),
            inp["instance_box"].astype(np.float32), inp["instance_box_volume"].astype(np.float32),
            inp["wall_box"], inp["wall_volume"])


def rel_err(a, b):
    import numpy as np
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
