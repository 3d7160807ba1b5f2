import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # pragma: no cover
        return False


@pytest.fixture(params=["host", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    """"host": the Python layer of tncontract_b200 over tests/fake_tnb.py (NumPy
    stand-in of the C ABI, checks label algebra / sweep logic on CPU);
    "gpu": the real libtnb.so on cuda:0 (the parity tests proper)."""
    if request.param == "host":
        import fake_tnb
        fake_tnb.install(monkeypatch)
    else:
        if not _cuda():
            pytest.skip("no CUDA device")
    return request.param
