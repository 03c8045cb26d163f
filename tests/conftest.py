import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    """-> (inputs list, rating, sd0, record) as torch tensors / dict."""
    import numpy as np
    import torch
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    ins = []
    k = 0
    while f"in/{k}" in z:
        ins.append(torch.from_numpy(z[f"in/{k}"]))
        k += 1
    sd0 = {key[3:]: torch.from_numpy(z[key]) for key in z.files if key.startswith("sd/")}
    return ins, torch.from_numpy(z["rating"]), sd0, z
