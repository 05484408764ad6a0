import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_spec_json(name):
    with open(os.path.join(GOLDEN, f"spec_{name}.json")) as f:
        spec = json.load(f)
    for g in spec:
        g["key"] = tuple(g["key"])
        g["state"] = [tuple(t) for t in g["state"]]
        g["node"] = [tuple(t) for t in g["node"]]
    return spec


def key_str(k):
    key, axis = k  # tuple or Axis
    return f"{key}:{axis}"


@pytest.fixture(scope="session")
def tiny_golden():
    import torch

    return torch.load(os.path.join(GOLDEN, "tiny_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def rn18_golden():
    import torch

    return torch.load(os.path.join(GOLDEN, "rn18_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def eval_golden():
    import torch

    return torch.load(os.path.join(GOLDEN, "eval_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def lap_golden():
    import numpy as np

    z = np.load(os.path.join(GOLDEN, "lap_golden.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return {n: {f: z[f"{n}/{f}"] for f in ("A", "maximize", "col", "obj")} for n in names}


@pytest.fixture(autouse=True)
def _exact_fp32_library_math():
    """cuDNN/cuBLAS TF32 is on by default for convolutions and would move activations by ~1e-3
    relative to the CPU oracle (SURVEY.md §7 'Forward passes'); parity tests compare exact-fp32
    forwards."""
    import torch

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
