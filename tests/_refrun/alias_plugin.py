"""pytest plugin used by tests/test_reference_own_tests.py: makes ``import refiners...`` resolve to the refiners_b200
mirror (every sub-module is registered under both names) and provides the fixtures of the reference's
tests/conftest.py that its weight-free unit tests need (that conftest itself cannot be imported: it pulls in
refiners.conversion, which needs diffusers / segment_anything - SURVEY.md section 8c)."""

import importlib
import pkgutil
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import refiners_b200  # noqa: E402


def _alias_all() -> None:
    sys.modules["refiners"] = refiners_b200
    for info in pkgutil.walk_packages(refiners_b200.__path__, prefix="refiners_b200."):
        if ".csrc" in info.name:
            continue
        try:
            mod = importlib.import_module(info.name)
        except Exception:  # optional sub-modules (PIL-dependent helpers ...) simply stay unaliased
            continue
        sys.modules["refiners" + info.name[len("refiners_b200"):]] = mod
    for name, mod in list(sys.modules.items()):  # modules registered at import time (reference-layout views, see layers/_paths.py)
        if name.startswith("refiners_b200."):
            sys.modules.setdefault("refiners" + name[len("refiners_b200"):], mod)


_alias_all()


@pytest.fixture(scope="session")
def test_device() -> torch.device:
    return torch.device("cpu")


@pytest.fixture(scope="session")
def test_device_zero() -> torch.device:
    return torch.device("cpu")


@pytest.fixture(scope="session", params=["float32"])
def test_dtype(request) -> torch.dtype:
    return getattr(torch, request.param)


@pytest.fixture(scope="session", params=["float32"])
def test_dtype_fp32_bf16_fp16(request) -> torch.dtype:
    return getattr(torch, request.param)


@pytest.fixture(scope="session", params=["float32"])
def test_dtype_fp32_fp16(request) -> torch.dtype:
    return getattr(torch, request.param)


@pytest.fixture(scope="session", params=["float32"])
def test_dtype_fp32_bf16(request) -> torch.dtype:
    return getattr(torch, request.param)


@pytest.fixture(scope="session", params=["float32"])
def test_dtype_fp16_bf16(request) -> torch.dtype:
    return getattr(torch, request.param)
