import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    try:  # fp32 torch comparators on the GPU must be true fp32 (TF32 silently changes the reference)
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    except Exception:
        pass
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_kat.npz"))


@pytest.fixture(scope="session")
def weight_digests():
    out = {}
    with open(os.path.join(ROOT, "tests", "golden", "weights_sha256.txt")) as f:
        for line in f:
            k, v = line.split()
            out[k] = v
    return out
