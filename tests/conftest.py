import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


SCENES_DIR = os.path.join(ROOT, "oracle", "_ref", "scenes")


def scene_blob(name):
    p = os.path.join(SCENES_DIR, name + ".bin")
    if not os.path.exists(p):
        pytest.skip("scene blob %s not built (oracle/make_scenes.py needs /root/reference)" % p)
    return p
