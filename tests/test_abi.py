"""CPU: the C-ABI library builds, loads and exports every symbol include/b200track.h declares;
the package refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import alufe_b200  # noqa: E402
from alufe_b200 import _lib  # noqa: E402


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200track.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    _lib.build()
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 27
    for n in names:
        assert hasattr(h, n), "libb200track.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "python bindings and header disagree"
    assert _lib.lib().b200_version() == 100


def test_argument_errors_without_gpu():
    lib = _lib.lib()
    rc = lib.b200_roi_align_fwd_f32(None, 7, 1, 1, 1, 1, None, 0, 1, 1, 1.0, 2, 1, None, None)
    assert rc == _lib.EINVAL and b"layout" in lib.b200_last_error()
    rc = lib.b200_roi_align_fwd_f32(None, 0, 1, 1, 1, 1, None, 0, 10, 10, 1.0, 2, 1, None, None)
    assert rc == _lib.OK                      # K == 0 is a no-op
    with pytest.raises(ValueError):
        _lib.check(lib.b200_lsap_f32(None, 1, 0, 3, 3, 1, 1.0, None, None, None, None))
    assert lib.b200_tracker_result_stride(None) == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.B200Error):
        alufe_b200.roi_align(torch.zeros(1, 4, 5, 5), torch.zeros(1, 5), (2, 2))
    with pytest.raises(_lib.B200Error):
        alufe_b200.hungarian_assign(np.ones((2, 2), np.float32))
    with pytest.raises(_lib.B200Error):
        alufe_b200.Tracking(conf=alufe_b200.SHIPPED_CONF)
    with pytest.raises(_lib.B200Error):
        alufe_b200.bbox_cost([[0, 0, 1, 1]], [[0, 0, 1, 1]], (1, 1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "a-lightweight-unsupervised-feature-extractor-_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
