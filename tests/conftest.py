import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


# ---- seeded synthetic inputs ------------------------------------------------------------
def clouds_uniform(rng, *shape):
    """U: uniform in [-1,1]^3."""
    return rng.uniform(-1.0, 1.0, size=shape).astype(np.float32)


def clouds_sphere(rng, *shape):
    """S: unit sphere + N(0, 0.01^2) radial noise (surface-like, small NN distances)."""
    v = rng.standard_normal(size=shape)
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v *= 1.0 + 0.01 * rng.standard_normal(size=shape[:-1] + (1,))
    return v.astype(np.float32)


def clouds_ties(rng, *shape):
    """T: coordinates quantised to multiples of 1/64 with 5 % exact duplicates (tie / adversarial set)."""
    v = np.round(rng.uniform(-1.0, 1.0, size=shape) * 64.0) / 64.0
    flat = v.reshape(-1, shape[-2], shape[-1])
    for cl in flat:
        n = cl.shape[0]
        dup = rng.choice(n, size=max(1, n // 20), replace=False)
        src = rng.choice(n, size=dup.size)
        cl[dup] = cl[src]
    return flat.reshape(shape).astype(np.float32)
