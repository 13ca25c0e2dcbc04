import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """-> (meta dict, {array name: ndarray}) for tests/golden/<name>.npz (oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["__meta__"]))
    return meta, {k: z[k] for k in z.files if k != "__meta__"}


def golden_sample(a, big=4096, n=257):
    """Same sub-sampling as oracle/make_golden.py:sample_of."""
    a = np.asarray(a).reshape(-1)
    if a.size <= big:
        return a
    idx = (np.arange(n, dtype=np.int64) * 7919) % a.size
    return a[idx]


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
