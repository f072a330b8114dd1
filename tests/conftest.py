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


def rel_l2(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def assert_close_flips(a, b, what, rtol=1e-4, atol=1e-5, l2=2e-5, outliers=0.0, outlier_atol=2e-2):
    """Element-wise fp32 tolerance + a norm-wise bound.  ``outliers``: fraction of elements allowed OUTSIDE the element-wise
    tolerance (but within ``outlier_atol`` of the tensor's scale).  Used for gradients only: an activation whose
    pre-activation sits within fp32 rounding of 0 (|z| < ~1e-7 |terms|) takes the other branch of (Leaky)ReLU / max-pool in
    one of the two evaluation orders; with 10^7..10^8 such units per test a handful flip, and each flip moves the gradient
    of its few fan-in entries by a visible amount while everything else agrees to rounding.  The norm-wise bound still
    holds the whole tensor to 1e-4."""
    ad, bd = a.detach().cpu().double(), b.detach().cpu().double()
    assert ad.shape == bd.shape, "%s: shape %s vs %s" % (what, tuple(ad.shape), tuple(bd.shape))
    scale = max(float(bd.abs().max()), 1e-30) if bd.numel() else 1.0
    err = (ad - bd).abs()
    bad = err > atol * max(scale, 1.0) + rtol * bd.abs()
    n_bad = int(bad.sum())
    allowed = max(1, int(outliers * ad.numel())) if outliers > 0 else 0      # a flip also moves whole-tensor sums (biases)
    assert n_bad <= allowed, "%s: %d/%d mismatches (allowed %d), max abs err %.3e (ref scale %.3e)" % (
        what, n_bad, ad.numel(), allowed, float(err.max()), scale)
    assert float(err.max()) <= outlier_atol * max(scale, 1.0) or n_bad == 0, "%s: outlier of %.3e (ref scale %.3e)" % (
        what, float(err.max()), scale)
    e = rel_l2(a, b)
    assert e <= l2, "%s: relative L2 error %.3e > %.1e" % (what, e, l2)
