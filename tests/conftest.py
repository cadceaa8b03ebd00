import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped automatically when no CUDA device is visible."""
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
def gold_dir():
    return GOLD


@pytest.fixture(scope="session")
def shipped_luts():
    """The reference's shipped x4 tables (tests/golden/luts_x4, copied by oracle/make_golden.py)."""
    luts = {}
    for s in (1, 2):
        for m in "sdy":
            luts["s{}_{}".format(s, m)] = np.load(
                os.path.join(GOLD, "luts_x4", "LUT_ft_x4_4bit_int8_s{}_{}.npy".format(s, m))
            ).reshape(-1, 1 if s == 1 else 16)
    return luts


@pytest.fixture(scope="session")
def set5():
    return dict(np.load(os.path.join(GOLD, "set5_x4.npz")))


@pytest.fixture(scope="session")
def pipeline_cases():
    data = np.load(os.path.join(GOLD, "ref_pipeline_cases.npz"))
    meta = json.load(open(os.path.join(GOLD, "ref_pipeline_cases.json")))
    return meta, data


@pytest.fixture(scope="session")
def interval_cases():
    """Reference-generated whole-pipeline cases at --interval 3/5/6/7 and scale 3 (oracle/make_golden.py)."""
    data = np.load(os.path.join(GOLD, "ref_interval_cases.npz"))
    meta = json.load(open(os.path.join(GOLD, "ref_interval_cases.json")))
    return meta, data


@pytest.fixture(scope="session")
def pass_cases():
    data = np.load(os.path.join(GOLD, "ref_pass_cases.npz"))
    meta = json.load(open(os.path.join(GOLD, "ref_pass_cases.json")))
    return meta, data


@pytest.fixture(scope="session")
def finetune_cases():
    data = np.load(os.path.join(GOLD, "ref_finetune_cases.npz"))
    meta = json.load(open(os.path.join(GOLD, "ref_finetune_cases.json")))
    return meta, data


def densify(idx, val, shape):
    a = np.zeros(int(np.prod(shape)), dtype=val.dtype)
    a[idx] = val
    return a.reshape(shape)


def norm_max_err(a, b):
    """max|a-b| / max|b|: the gradient-parity metric (SURVEY.md 8-SPEC noise floor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())
