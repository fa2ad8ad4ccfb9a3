import os
import sys

import numpy as np
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
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_model(name):
    """-> (cfg dict, state_dict of torch tensors, raw npz)"""
    z = load_golden(f"model_{name}.npz")
    cfg = {}
    for k in z.files:
        if k.startswith("cfg_"):
            v = z[k]
            cfg[k[4:]] = v.item() if v.dtype.kind in "iub" else str(v)
    cfg["causal"] = bool(cfg["causal"])
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    return cfg, sd, z


def rel_err(a, b):
    """max |a-b| / max |b|  — the 'max-rel-err' every tolerance in this suite is stated in."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 — the per-tensor gradient metric (see tests/golden/make_golden_fp64.py for why not max-norm)."""
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def grad_tolerance(ref32_err):
    """1e-3 (north star), or 3x the fp32 reference's own error against fp64 truth where that is larger."""
    return max(1e-3, 3.0 * ref32_err)


def assert_gradients_match(pairs):
    """pairs: (name, err_mine, err_ref32), relative L2 against fp64 truth per tensor.

    fp32 gradients of this network carry isolated outliers: one pre-activation within rounding distance of a PReLU kink
    that lands on the other side changes a whole row of a weight gradient by ~1/sqrt(frames) (6e-3 at 24k frames).  The
    reference's own fp32 autograd shows them too (e.g. 3e-3 on one gamma of the softmax config).  So the bar is:
    the typical tensor is as accurate as the reference's, every tensor is within max(1e-3, 3x reference) except for
    at most 2 % outliers, and no tensor is off by more than 2e-2."""
    import statistics
    errs = [e for _, e, _ in pairs]
    refs = [r for _, _, r in pairs]
    assert statistics.median(errs) <= max(1e-4, 2.0 * statistics.median(refs)), (statistics.median(errs), statistics.median(refs))
    bad = [(k, e, r) for k, e, r in pairs if e > grad_tolerance(r)]
    assert len(bad) <= max(1, int(0.02 * len(pairs))), bad
    assert max(errs) < 2e-2, max(pairs, key=lambda t: t[1])
