"""Oracle (test-only): builds and loads the C restatements under oracle/csrc with gcc.

``build()`` is called by ``__graft_entry__.build()`` and lazily on first use; the
resulting ``oracle/_build/liboracle.so`` is git-ignored and travels to the GPU box.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = [os.path.join(_HERE, "csrc", n) for n in ("lsap_ref.c", "roi_align_ref.c")]
_OUT = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(_OUT) and all(
            os.path.getmtime(_OUT) >= os.path.getmtime(s) for s in _SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _OUT] + _SRC + ["-lm"]
    subprocess.run(cmd, check=True)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_lsap_f64.restype = ctypes.c_int
        _lib.oracle_roi_align_f32.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def lsap(cost):
    """Returns (col_of_row[int64 nr], row_of_col[int64 nc]); raises ValueError like scipy."""
    c = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = c.shape
    c4r = np.empty(nr, dtype=np.int64)
    r4c = np.empty(nc, dtype=np.int64)
    rc = lib().oracle_lsap_f64(ctypes.c_int64(nr), ctypes.c_int64(nc), _p(c), _p(c4r), _p(r4c))
    if rc == -1:
        raise ValueError("matrix contains invalid numeric entries")
    if rc == -2:
        raise ValueError("cost matrix is infeasible")
    return c4r, r4c


def roi_align(feat, rois, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    f = np.ascontiguousarray(feat, dtype=np.float32)
    r = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    B, C, H, W = f.shape
    PH, PW = output_size
    out = np.empty((r.shape[0], C, PH, PW), dtype=np.float32)
    rc = lib().oracle_roi_align_f32(_p(f), B, C, H, W, _p(r), ctypes.c_int64(r.shape[0]), PH, PW,
                                    ctypes.c_float(spatial_scale), int(sampling_ratio),
                                    int(bool(aligned)), _p(out))
    if rc != 0:
        raise ValueError("roi batch index out of range")
    return out
