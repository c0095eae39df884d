import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def voc(golden):
    v = golden("voc_flat.npz")
    return {k: (v[k] if v[k].ndim else v[k].item()) for k in v.files}


@pytest.fixture(scope="session")
def scene_frames():
    """A few full-size (1241x376) synthetic stereo pairs, rendered once per session."""
    from srba_stereo_slam_b200 import synth
    cache = {}

    def get(n, width=1241, height=376, seed=0):
        key = (n, width, height, seed)
        if key not in cache:
            cache[key] = synth.CorridorScene(seed=seed, width=width, height=height).sequence(n)
        return cache[key]
    return get
