"""Cost operators with the reference's names (model/utils/costTool/costCard.py), on the GPU.

Inputs are the reference's host-side Python lists (or tensors); results are CUDA float32 tensors.
"""
from typing import Any, Dict, List, Optional, Tuple

import torch

from . import _lib


def _dev(device=None):
    if not torch.cuda.is_available():
        raise _lib.B200Error("no CUDA device: this package has no CPU path")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _f32(x, device, cols=None):
    t = x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=torch.float32)
    t = t.to(device=device, dtype=torch.float32)
    if cols is not None:
        t = t.reshape(-1, cols)
    return t.contiguous()


def _pair(C_app, boxes_prev, boxes_cur, conf_prev, conf_cur, w, want, device):
    bp, bc = _f32(boxes_prev, device, 4), _f32(boxes_cur, device, 4)
    M, N = bp.shape[0], bc.shape[0]
    cp = _f32(conf_prev, device).reshape(-1) if conf_prev is not None else torch.ones(M, device=device)
    cc = _f32(conf_cur, device).reshape(-1) if conf_cur is not None else torch.ones(N, device=device)
    out = {k: torch.zeros((M, N), dtype=torch.float32, device=device) for k in want}
    if M == 0 or N == 0:
        return out
    if C_app is not None:
        C_app = _f32(C_app, device).reshape(M, N)
    with torch.cuda.device(device):
        rc = _lib.lib().b200_pair_cost_f32(
            _lib.ptr(C_app), _lib.ptr(bp), _lib.ptr(bc), _lib.ptr(cp), _lib.ptr(cc), M, N,
            w["w_app"], w["w_bbox"], w["w_conf"], w["alpha"], w["beta"], w["eps"],
            _lib.ptr(out.get("C_total")), _lib.ptr(out.get("C_bbox")), _lib.ptr(out.get("C_center")),
            _lib.ptr(out.get("C_scale")), _lib.ptr(out.get("C_conf")), N, _lib.stream_ptr(device))
    _lib.check(rc)
    return out


_W0 = dict(w_app=1.0, w_bbox=0.0, w_conf=0.0, alpha=1.0, beta=1.0, eps=1e-6)


@torch.no_grad()
def bbox_cost(boxes_prev: List[List[float]], boxes_cur: List[List[float]], input_hw: Tuple[int, int],
              alpha: float = 1.0, beta: float = 1.0) -> Dict[str, torch.Tensor]:
    """costCard.py:109-174.  ``input_hw`` is accepted and unused, as in the reference (:147-149)."""
    w = dict(_W0, alpha=float(alpha), beta=float(beta))
    return _pair(None, boxes_prev, boxes_cur, None, None, w, ("C_center", "C_scale", "C_bbox"), _dev())


@torch.no_grad()
def conf_cost(conf_prev: List[float], conf_cur: List[float], eps: float = 1e-6) -> torch.Tensor:
    """costCard.py:178-203."""
    dev = _dev()
    M, N = len(conf_prev), len(conf_cur)
    zp, zc = torch.zeros((M, 4)), torch.zeros((N, 4))
    return _pair(None, zp, zc, conf_prev, conf_cur, dict(_W0, eps=float(eps)), ("C_conf",), dev)["C_conf"]


@torch.no_grad()
def cal_cost(*, C_app: torch.Tensor, boxes_prev, boxes_cur, input_hw, conf_prev, conf_cur,
             w_app: float = 1.0, w_bbox: float = 0.3, w_conf: float = 0.2, alpha: float = 1.0,
             beta: float = 0.5, assign: Optional[List[int]] = None, unmatch_cost: float = 10.0) -> Dict[str, Any]:
    """costCard.py:206-300: one fused kernel for every term and the weighted sum."""
    dev = C_app.device if isinstance(C_app, torch.Tensor) and C_app.is_cuda else _dev()
    w = dict(w_app=float(w_app), w_bbox=float(w_bbox), w_conf=float(w_conf), alpha=float(alpha),
             beta=float(beta), eps=1e-6)
    out = _pair(C_app, boxes_prev, boxes_cur, conf_prev, conf_cur, w,
                ("C_total", "C_bbox", "C_center", "C_scale", "C_conf"), dev)
    out["C_app"] = _f32(C_app, dev).reshape(out["C_total"].shape)
    if assign is not None:                       # :282-298, host bookkeeping on a copy of C_total
        C = out["C_total"].detach().cpu().numpy()
        total, used = 0.0, set()
        for i, j in enumerate(assign):
            if j == -1:
                total += unmatch_cost
            elif j in used:
                total += 1e6
            else:
                total += C[i, j]
                used.add(j)
        out["total_cost"] = float(total)
    return out


@torch.no_grad()
def app_cost_topk(bank: torch.Tensor, bank_len: torch.Tensor, det: torch.Tensor, topk: int = 5,
                  use_topk_mean: bool = True, fallback: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched Tracking.build_C_app_topk (mainTracking.py:141-211).

    bank [M,T,128] float32, bank_len [M] int32 (valid rows), det [N,128]; ``fallback`` [M,128] is the
    EMA embedding used when a bank is empty (:180-182); without it an empty bank gives a row of ones.
    """
    _lib.require_cuda(bank, "bank")
    dev = bank.device
    M, T, D = bank.shape
    if D != 128 or det.shape[-1] != 128:
        raise ValueError("embeddings must be 128-D")
    N = det.shape[0]
    out = torch.zeros((M, N), dtype=torch.float32, device=dev)
    if M == 0 or N == 0:
        return out
    bank = bank.to(torch.float32).contiguous()
    det = det.to(device=dev, dtype=torch.float32).contiguous()
    bl = bank_len.to(device=dev, dtype=torch.int32).contiguous()
    fb = fallback.to(device=dev, dtype=torch.float32).contiguous() if fallback is not None else None
    with torch.cuda.device(dev):
        rc = _lib.lib().b200_app_cost_topk_f32(_lib.ptr(bank), _lib.ptr(bl), _lib.ptr(fb), _lib.ptr(det), M, N, T,
                                               int(topk), int(bool(use_topk_mean)), _lib.ptr(out), N,
                                               _lib.stream_ptr(dev))
    _lib.check(rc)
    return out
