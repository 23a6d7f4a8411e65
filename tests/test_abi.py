"""The C-ABI library loads without a GPU and exports exactly what include/gpk.h declares."""
import ctypes
import os
import re

import pytest

from pygps_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "gpk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpk_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libgpk.so does not export %s" % n


def test_binding_covers_the_header():
    assert sorted(p[0] for p in _lib.PROTOTYPES) == _declared()


def test_no_torch_or_python_types_in_signatures():
    text = open(os.path.join(ROOT, "include", "gpk.h")).read()
    assert 'extern "C"' in text
    decls = re.sub(r"/\*.*?\*/", "", text, flags=re.S)      # declarations only, comments stripped
    for banned in ("torch", "at::", "PyObject", "Tensor", "std::"):
        assert banned not in decls, banned


def test_version_and_errors_need_no_gpu(lib):
    assert lib.gpk_version() >= 100
    assert b"invalid argument" in lib.gpk_strerror(-1)
    assert b"positive definite" in lib.gpk_strerror(7)
    assert lib.gpk_create(0, None) == -1
    assert lib.gpk_destroy(None) == -1


def test_stats_struct_layout():
    assert ctypes.sizeof(_lib.GpkStats) == 7 * 8 + 3 * 8


def test_fails_loudly_without_cuda(lib):
    """No CPU fallback: on a box without a CUDA device the product path must raise, not compute."""
    cnt = ctypes.c_int(-1)
    rc = lib.gpk_device_count(ctypes.byref(cnt))
    if rc == 0 and cnt.value > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(_lib.GpkError):
        _lib.Engine()
    import numpy as np
    import pygps_b200 as pg
    with pytest.raises(_lib.GpkError):
        pg.GPR().getPosterior(np.zeros((4, 1)), np.ones((4, 1)))
    with pytest.raises(_lib.GpkError):
        pg.cov.RBF().getCovMatrix(x=np.zeros((4, 1)), mode='train')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pygps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "gp_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
