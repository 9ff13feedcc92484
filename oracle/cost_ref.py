"""Oracle (test-only): the reference's cost operators restated on numpy float32.

Follows model/utils/costTool/costCard.py (bbox_cost :109-174, conf_cost :178-203,
cal_cost :206-300) and Tracking.build_C_app_topk (model/mainTracking.py:141-211).
The reference evaluates these with torch CPU float32 tensors built from Python
floats; numpy float32 arrays follow the same operation order.
"""
import numpy as np

F32 = np.float32


def _f32(a, cols=None):
    arr = np.asarray(a, dtype=F32)
    if cols is not None:
        arr = arr.reshape(-1, cols)
    return arr


def bbox_cost(boxes_prev, boxes_cur, input_hw, alpha=1.0, beta=1.0):
    """costCard.py:109-174.  ``input_hw`` is accepted and unused, like the reference
    (the image diagonal is computed at :147-149 and never applied)."""
    M, N = len(boxes_prev), len(boxes_cur)
    if M == 0 or N == 0:
        z = np.zeros((M, N), dtype=F32)
        return {"C_center": z, "C_scale": z, "C_bbox": z}
    bp, bc = _f32(boxes_prev, 4), _f32(boxes_cur, 4)
    half = F32(0.5)
    cp = half * (bp[:, :2] + bp[:, 2:])
    cc = half * (bc[:, :2] + bc[:, 2:])
    d = cp[:, None, :] - cc[None, :, :]
    dist = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])
    wp = np.maximum(bp[:, 2] - bp[:, 0], F32(1.0))
    hp = np.maximum(bp[:, 3] - bp[:, 1], F32(1.0))
    scale_p = np.maximum(np.sqrt(wp * wp + hp * hp), F32(1.0))
    C_center = dist / scale_p[:, None]
    Ap = wp * hp
    wc = np.maximum(bc[:, 2] - bc[:, 0], F32(1.0))
    hc = np.maximum(bc[:, 3] - bc[:, 1], F32(1.0))
    Ac = wc * hc
    C_scale = np.abs(np.log(np.maximum(Ac[None, :] / Ap[:, None], F32(1e-6))))
    C_bbox = F32(alpha) * C_center + F32(beta) * C_scale
    return {"C_center": C_center, "C_scale": C_scale, "C_bbox": C_bbox}


def conf_cost(conf_prev, conf_cur, eps=1e-6):
    """costCard.py:178-203.  |log(max(cc,eps)/max(cp,eps))|."""
    M, N = len(conf_prev), len(conf_cur)
    if M == 0 or N == 0:
        return np.zeros((M, N), dtype=F32)
    cp = np.maximum(_f32(conf_prev), F32(eps))
    cc = np.maximum(_f32(conf_cur), F32(eps))
    return np.abs(np.log(cc[None, :] / cp[:, None]))


def cal_cost(*, C_app, boxes_prev, boxes_cur, input_hw, conf_prev, conf_cur,
             w_app=1.0, w_bbox=0.3, w_conf=0.2, alpha=1.0, beta=0.5,
             assign=None, unmatch_cost=10.0):
    """costCard.py:206-300.  Weighted sum + optional scalar fitness of an assignment."""
    C_app = np.asarray(C_app, dtype=F32)
    bb = bbox_cost(boxes_prev, boxes_cur, input_hw, alpha=alpha, beta=beta)
    C_conf = conf_cost(conf_prev, conf_cur)
    C_total = F32(w_app) * C_app + F32(w_bbox) * bb["C_bbox"] + F32(w_conf) * C_conf
    out = {"C_total": C_total, "C_app": C_app, "C_bbox": bb["C_bbox"],
           "C_center": bb["C_center"], "C_scale": bb["C_scale"], "C_conf": C_conf}
    if assign is not None:
        total, taken = 0.0, set()
        for i, j in enumerate(assign):
            if j == -1:
                total += unmatch_cost
            elif j in taken:
                total += 1e6
            else:
                total += C_total[i, j]
                taken.add(j)
        out["total_cost"] = float(total)
    return out


def _unit_rows(a):
    return a / (np.linalg.norm(a, axis=1, keepdims=True) + 1e-12)


def app_cost_topk(banks, det_embs, topk=5, use_topk_mean=True, fallback_embs=None):
    """Tracking.build_C_app_topk, mainTracking.py:141-211.

    banks: list (len M) of lists of 128-D vectors (a track's feat_historical);
    fallback_embs: per-track EMA embedding used when the bank is empty (or None).
    Returns float32 [M, N] = 1 - mean(top-k over the bank of cos(bank_t, det_j)).
    """
    M, N = len(banks), len(det_embs)
    if M == 0 or N == 0:
        return np.zeros((M, N), dtype=F32)
    det = _unit_rows(np.stack([np.asarray(e, dtype=F32).reshape(-1) for e in det_embs]))
    rows = []
    for i, bank in enumerate(banks):
        if bank is None or len(bank) == 0:
            fb = None if fallback_embs is None else fallback_embs[i]
            if fb is None:
                rows.append(np.ones((N,), dtype=F32))
                continue
            bank = [fb]
        B = _unit_rows(np.stack([np.asarray(f, dtype=F32).reshape(-1) for f in bank]))
        sim = B @ det.T                                   # [T, N]
        k = min(int(topk), sim.shape[0])
        if k <= 0:
            rows.append(np.ones((N,), dtype=F32))
            continue
        if use_topk_mean:
            top = -np.sort(-sim, axis=0)[:k]
            s = top.mean(axis=0, dtype=F32)
        else:
            s = sim.max(axis=0)
        rows.append((F32(1.0) - s).astype(F32))
    return np.stack(rows, axis=0)
