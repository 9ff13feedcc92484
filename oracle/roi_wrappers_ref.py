"""Oracle (test-only): the reference's two ROI Align call conventions restated on numpy + the C ROI oracle.

* ``roi_align_from_input_boxes`` - MainInfer.roi_align_from_input_boxes, tracking.py:193-221
  (= tracking_win.py:239-267, model/utils/inferScr/infer.py:143-170).
* ``preprocess_roi``             - PreProcess._preprocess_roi, model/utils/trainingScr/trainingCard.py:24-79.

Pinned by tests/golden/roi_wrappers.npz, which tests/golden/make_golden.py records by calling the unmodified
reference methods unbound (both ignore ``self``).
"""
import numpy as np

from . import native


def roi_align_from_input_boxes(feat, boxes_in, input_hw, out_size=(7, 7), aligned=True, sampling_ratio=2):
    f = np.asarray(feat, dtype=np.float32)
    H_in, _ = input_hw                                           # tracking.py:205: W_in is unpacked and unused
    Hf = f.shape[2]
    scale = Hf / float(H_in)                                     # :207
    b = np.asarray(boxes_in, dtype=np.float32).reshape(-1, 4)    # :209-213 rois in the map's dtype, batch index 0
    rois = np.concatenate([np.zeros((b.shape[0], 1), np.float32), b], axis=1)
    return native.roi_align(f, rois, tuple(out_size), scale, sampling_ratio, aligned)


def preprocess_rois(boxes_xyxy, feat_hw, img_hw, enforce_min_size=1.0):
    """trainingCard.py:37-68: the [N,5] float32 roi tensor the reference hands to roi_align (float32 arithmetic
    step by step like the torch code)."""
    f32 = np.float32
    b = np.asarray(boxes_xyxy, dtype=f32).reshape(-1, 4)
    Hf, Wf = feat_hw
    img_h, img_w = img_hw
    x1, x2 = np.minimum(b[:, 0], b[:, 2]), np.maximum(b[:, 0], b[:, 2])          # :45-49
    y1, y2 = np.minimum(b[:, 1], b[:, 3]), np.maximum(b[:, 1], b[:, 3])
    sx, sy = f32(Wf / float(img_w)), f32(Hf / float(img_h))                       # :52-57 (python float -> float32 mul)
    x1, x2, y1, y2 = x1 * sx, x2 * sx, y1 * sy, y2 * sy
    x1, x2 = np.clip(x1, 0, Wf - 1).astype(f32), np.clip(x2, 0, Wf - 1).astype(f32)   # :59-62
    y1, y2 = np.clip(y1, 0, Hf - 1).astype(f32), np.clip(y2, 0, Hf - 1).astype(f32)
    if enforce_min_size > 0:                                                      # :64-68
        x2 = np.clip(np.maximum(x2, x1 + f32(enforce_min_size)), 0, Wf - 1).astype(f32)
        y2 = np.clip(np.maximum(y2, y1 + f32(enforce_min_size)), 0, Hf - 1).astype(f32)
    return np.stack([np.zeros_like(x1), x1, y1, x2, y2], axis=1).astype(f32)


def preprocess_roi(feat, bboxes_xyxy, img_hw, output_size=(10, 10), sampling_ratio=2, aligned=True,
                   enforce_min_size=1.0):
    f = np.asarray(feat, dtype=np.float32)
    assert f.ndim == 4 and f.shape[0] == 1, "feat shape expected [1,C,H,W], got %s" % (f.shape,)   # :34
    rois = preprocess_rois(bboxes_xyxy, f.shape[2:], img_hw, enforce_min_size)
    return native.roi_align(f, rois, tuple(output_size), 1.0, sampling_ratio, aligned)           # :71-78
