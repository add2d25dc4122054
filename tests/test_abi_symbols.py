"""The C-ABI library builds for sm_100a, loads without a GPU, exports every
symbol include/argweaver_b200.h declares, and fails loudly (no CPU fallback)
when no CUDA device is present."""

import ctypes
import os
import re

import pytest

from argweaver_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "argweaver_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    skip = {"defined", "sizeof"}
    return sorted(set(n for n in names if n not in skip))


def test_library_exports_declared_symbols():
    so = build.build_cuda()
    lib = ctypes.CDLL(so)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.AwbError):
        api.Context(0)
