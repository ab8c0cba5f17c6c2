import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nns-cuda_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.build()
    return o


@pytest.fixture(scope="session")
def nns():
    import nns_b200

    return nns_b200


@pytest.fixture(scope="session")
def datagen():
    from nns_b200 import datagen as d

    return d


def make_case(kind: str, k: int, m: int, n: int, seed: int):
    """Seeded test inputs shared by CPU and GPU tests."""
    from nns_b200 import datagen as d

    if kind == "uniform":
        return d.uniform_points(m, k, seed, 0), d.uniform_points(n, k, seed, 1)
    if kind == "grid":  # coarse grid => many exact ties and duplicates
        s = np.floor(d.uniform_points(m, k, seed, 0) * 8) / 8
        r = np.floor(d.uniform_points(n, k, seed, 1) * 8) / 8
        return s.astype(np.float32), r.astype(np.float32)
    if kind == "clustered":
        return d.clustered_workload(m, n, k, seed)
    raise ValueError(kind)
