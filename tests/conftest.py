import json
import os
import shutil
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from _pkgload import load_package  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package.  Builds libgrl_b200.so first when a compiler is present and the library is stale
    (the GPU box receives the prebuilt .so with the snapshot and never needs to)."""
    p = load_package()
    from grl_image_restoration_b200 import build as b

    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        b.build()
    return p


@pytest.fixture(scope="session")
def oracle():
    import grl_oracle

    return grl_oracle


@pytest.fixture(scope="session")
def cases():
    with open(os.path.join(GOLD, "cases.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def geometry_golden():
    with open(os.path.join(GOLD, "geometry.json")) as f:
        return json.load(f)


def load_npz(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLD, name)).items()}


@pytest.fixture(scope="session")
def golden_loader():
    return load_npz


@pytest.fixture(scope="session")
def device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
