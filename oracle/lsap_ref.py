"""Oracle (test-only): hungarian_assign restated (model/utils/costTool/hung.py:5-45).

The solver is oracle/csrc/lsap_ref.c (scipy's algorithm restated); the wrapper keeps
hung.py's contract: empty-shape shortcuts (:19-25), solve on the full matrix (:28),
keep pairs with C[i,j] <= cost_max (:35-40), unmatched lists ascending (:42-43).
"""
import numpy as np

from . import native


def linear_sum_assignment(cost):
    """(row_ind, col_ind) like scipy: rows ascending, min(nr,nc) pairs."""
    c4r, _ = native.lsap(cost)
    rows = np.nonzero(c4r >= 0)[0]
    return rows.astype(np.int64), c4r[rows]


def hungarian_assign(C_total, cost_max=1e9):
    C = np.asarray(C_total)
    M, N = C.shape
    if M == 0 and N == 0:
        return [], [], []
    if M == 0:
        return [], [], list(range(N))
    if N == 0:
        return [], list(range(M)), []
    rows, cols = linear_sum_assignment(C)
    matches = [(int(i), int(j)) for i, j in zip(rows, cols) if float(C[i, j]) <= float(cost_max)]
    got_r = {i for i, _ in matches}
    got_c = {j for _, j in matches}
    return (matches, [i for i in range(M) if i not in got_r],
            [j for j in range(N) if j not in got_c])
