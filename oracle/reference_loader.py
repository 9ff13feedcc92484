"""Import the UNMODIFIED reference from /root/reference -- build container only.

Used by tests/golden/make_golden.py (to generate fixtures) and by the few
``-m "not gpu"`` tests that pin the oracle restatements against the live reference
when it is mounted.  /root/reference does not exist on the GPU box: nothing on the
``-m gpu`` / smoke / bench paths may call this.

Two stubs are needed: the filterpy shim, and a dummy
``model.utils.inferScr.infer`` because mainTracking.py:1 imports an unused symbol
that drags in matplotlib/seaborn.  ``Tracking.__init__`` reads a CWD-relative YAML
(mainTracking.py:47), so the constructor helper chdirs for the duration of the call.
"""
import contextlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200TRACK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "mainTracking.py"))


def _install_stubs():
    from . import filterpy_shim
    filterpy_shim.install()
    name = "model.utils.inferScr.infer"
    if name not in sys.modules:
        stub = types.ModuleType(name)
        stub.MainInfer = type("MainInfer", (), {})
        sys.modules[name] = stub
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load():
    """Returns a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    import model.mainTracking as mt
    import model.utils.costTool.costCard as cc
    import model.utils.costTool.KalmanFilter as kf
    import model.utils.costTool.hung as hg
    return types.SimpleNamespace(mainTracking=mt, costCard=cc, KalmanFilter=kf, hung=hg)


def new_tracking():
    ref = load()
    with _cwd(REFERENCE_ROOT):
        return ref.mainTracking.Tracking()


def load_roi_wrappers():
    """The reference's two ROI Align call conventions as unbound functions (both ignore ``self``):
    ``MainInfer.roi_align_from_input_boxes`` (tracking.py:193-221) and ``PreProcess._preprocess_roi``
    (trainingCard.py:24-79).  trainingCard.py:9 imports the YOLOv7 wrapper, which drags in matplotlib /
    seaborn (absent): a dummy ``model.yolov7.yoloDetects2`` stands in, it is never called."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_stubs()
    import model  # noqa: F401  (the reference's top-level package)
    for name in ("model.yolov7", "model.yolov7.yoloDetects2"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.__path__ = []
            sys.modules[name] = stub
    sys.modules["model.yolov7.yoloDetects2"].YoloDetects = type("YoloDetects", (), {})
    sys.modules["model.yolov7"].yoloDetects2 = sys.modules["model.yolov7.yoloDetects2"]
    sys.modules["model"].yolov7 = sys.modules["model.yolov7"]
    import tracking as live
    import model.utils.trainingScr.trainingCard as tc
    return types.SimpleNamespace(roi_align_from_input_boxes=live.MainInfer.roi_align_from_input_boxes,
                                 preprocess_roi=tc.PreProcess._preprocess_roi)
