"""hungarian_assign (model/utils/costTool/hung.py:5-45) on the GPU."""
from typing import List, Tuple

import numpy as np
import torch

from . import _lib

_ERR = {_lib.ENUMERIC: "matrix contains invalid numeric entries", _lib.EINFEASIBLE: "cost matrix is infeasible"}


def lsap_batched(C: torch.Tensor, cost_max: float = 1e9):
    """C: CUDA float32 [B,M,N].  Returns (col_of_row int32 [B,M], matched uint8 [B,M], status int32 [B])."""
    _lib.require_cuda(C, "C")
    if C.dim() != 3:
        raise ValueError("C must be [B, M, N]")
    C = C.to(torch.float32).contiguous()
    B, M, N = C.shape
    col = torch.full((B, M), -1, dtype=torch.int32, device=C.device)
    ok = torch.zeros((B, M), dtype=torch.uint8, device=C.device)
    status = torch.zeros((B,), dtype=torch.int32, device=C.device)
    with torch.cuda.device(C.device):
        rc = _lib.lib().b200_lsap_f32(_lib.ptr(C), B, M * N, M, N, max(N, 1), float(cost_max), _lib.ptr(col),
                                      _lib.ptr(ok), _lib.ptr(status), _lib.stream_ptr(C.device))
    _lib.check(rc)
    return col, ok, status


def hungarian_assign(C_total, cost_max: float = 1e9) -> Tuple[List[Tuple[int, int]], List[int], List[int]]:
    """Same contract as hung.py: (matches by ascending row, unmatched rows, unmatched columns).
    ``C_total`` may be a numpy array or a (CPU/CUDA) tensor; the solve runs on the current CUDA device.
    Raises ValueError for NaN/-inf entries or an infeasible matrix, like scipy does."""
    if isinstance(C_total, torch.Tensor):
        C = C_total.detach()
    else:
        C = torch.from_numpy(np.ascontiguousarray(np.asarray(C_total), dtype=np.float32))
    if C.dim() != 2:
        raise ValueError("expected a matrix (2-D array), got a %d array" % C.dim())
    M, N = C.shape
    if M == 0 and N == 0:
        return [], [], []
    if M == 0:
        return [], [], list(range(N))
    if N == 0:
        return [], list(range(M)), []
    if not C.is_cuda:
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device: this package has no CPU path")
        C = C.cuda()
    col, ok, status = lsap_batched(C.reshape(1, M, N), cost_max)
    col, ok, st = col[0].cpu().numpy(), ok[0].cpu().numpy(), int(status[0].item())
    if st != 0:
        raise ValueError(_ERR.get(st, "assignment failed (%d)" % st))
    matches = [(int(i), int(col[i])) for i in range(M) if ok[i]]
    used = {j for _, j in matches}
    return matches, [i for i in range(M) if not ok[i]], [j for j in range(N) if j not in used]
