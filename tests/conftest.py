import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLD, "meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def thresholds():
    with open(os.path.join(GOLD, "opt_thresholds.json")) as f:
        return json.load(f)


_SD_CACHE = {}


def synthetic_sd(model_type, sr=16000):
    from sed_b200 import synth
    key = (model_type, sr)
    if key not in _SD_CACHE:
        _SD_CACHE[key] = synth.synthetic_state_dict(model_type, sr)
    return _SD_CACHE[key]


def logmel_close(got, ref, rtol=1e-4):
    """north_star front-end tolerance: |a-b| <= 1e-4 * max(|b|, 1)  (SURVEY.md section 7 'hard parts')."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    tol = rtol * np.maximum(np.abs(ref), 1.0)
    return np.abs(got - ref) <= tol
