"""Shared pytest plumbing: the ``gpu`` marker, golden-fixture loading, comparison helpers."""
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


@pytest.fixture(scope="session", autouse=True)
def _native_library_is_current():
    """The .so is git-ignored; (re)build it in-tree when it is missing or older than its sources (a no-op
    otherwise: per-source hashes).  The product itself never builds or falls back -- it raises."""
    from peppa_b200 import build
    build.build()


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def from_bits(a):
    """uint16 bf16 bit patterns (as stored by oracle/make_golden.py) -> float32 tensor."""
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(torch.bfloat16).float()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype == np.uint16:
            out[k] = from_bits(v)
        elif v.dtype.kind in "fiu" and v.ndim > 0:
            out[k] = torch.from_numpy(v)
        else:
            out[k] = v
    return out


def golden_files(prefix):
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


def rel_err(x, ref):
    """Norm-wise relative error max|x-ref| / max|ref| (the 1e-3 loss/gradient bar of BASELINE.json)."""
    x = torch.as_tensor(x, dtype=torch.float64).cpu()
    ref = torch.as_tensor(ref, dtype=torch.float64).cpu()
    denom = ref.abs().max().clamp_min(1e-30)
    return ((x - ref).abs().max() / denom).item()


def row_rel_err(x, ref):
    """Row-wise relative error: max over rows i of ||x_i - ref_i||_2 / max(||ref_i||_2, 1% of the median row norm).
    Beside the norm-wise bar it keeps a few wrong gradient rows from hiding under one large entry elsewhere; the
    floor only protects rows whose true gradient is (nearly) zero."""
    x = torch.as_tensor(x, dtype=torch.float64).cpu()
    ref = torch.as_tensor(ref, dtype=torch.float64).cpu()
    if x.dim() < 2:
        return rel_err(x, ref)
    rn = ref.norm(dim=-1)
    floor = (0.01 * rn.median()).clamp_min(1e-30)
    return ((x - ref).norm(dim=-1) / torch.maximum(rn, floor)).max().item()
