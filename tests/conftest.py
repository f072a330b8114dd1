import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    return _cache[name]


def as_batch(fields, device=None):
    ns = types.SimpleNamespace()
    for k, v in fields.items():
        setattr(ns, k, v.to(device) if (device is not None and torch.is_tensor(v)) else v)
    return ns


def assert_close(a, b, rtol=1e-4, atol=1e-5, what=""):
    """rtol 1e-4 is the fp32 tolerance BASELINE.json's north_star states; atol scales with the data."""
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    scale = max(float(b.abs().max()), 1e-30) if b.numel() else 1.0
    err = (a - b).abs()
    tol = atol * max(scale, 1.0) + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), "%s: %d/%d mismatches, max abs err %.3e (ref scale %.3e)" % (
        what, int(bad.sum()), a.numel(), float(err.max()), scale)
