import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def load_pkg():
    """Import the product package (its directory name has hyphens, so it is loaded by path)."""
    if "sdso_b200" in sys.modules:
        return sys.modules["sdso_b200"]
    path = os.path.join(ROOT, "stereo-dso-g2o_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location("sdso_b200", path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sdso_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def scene():
    import synth
    return synth.make_scene()


@pytest.fixture(scope="session")
def frames(scene):
    """Rendered left images + depth for camera indices 0..3 and the right image of frame 0 (cached per session)."""
    import synth
    out = {}
    for k in range(3):
        out[k] = synth.render(scene, synth.camera_pose(k))
    out["r0"] = synth.render(scene, synth.right_of(synth.camera_pose(0)))
    return out
