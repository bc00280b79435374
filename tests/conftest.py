import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "otf_goldens.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the shared library is a build artefact (git-ignored): build it once if this checkout has none
    lib = os.path.join(ROOT, "trainner_redux_b200", "libotf_b200.so")
    if not os.path.exists(lib):
        import subprocess

        subprocess.run(["make", "-C", os.path.join(ROOT, "trainner_redux_b200", "csrc"), "-j8"], check=True)


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    z = np.load(GOLDEN)
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def dev():
    return torch.device("cuda:0")
