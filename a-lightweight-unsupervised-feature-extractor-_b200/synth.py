"""Seeded synthetic inputs for the tracker hot path.

Pure numpy; shared by tests, bench.py and tests/golden/make_golden.py so every arm
(GPU path, oracle, live reference) sees byte-identical inputs.
"""
import numpy as np

CONFIGS = {
    # name: (Hf, Wf, H_in, W_in, N boxes)
    "c1": (20, 20, 640, 640, 8),
    "c2": (40, 40, 1280, 1280, 64),
    "c4": (40, 40, 1280, 1280, 512),
    "c5": (34, 60, 1088, 1920, 128),
}


def feature_map(seed, B, C, Hf, Wf):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((B, C, Hf, Wf), dtype=np.float32)


def random_boxes(rng, n, H_in, W_in, lo=24.0, hi=96.0):
    cx = rng.uniform(0.1, 0.9, n) * W_in
    cy = rng.uniform(0.1, 0.9, n) * H_in
    w = rng.uniform(lo, hi, n)
    h = rng.uniform(lo, hi, n)
    return np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)


def edge_case_boxes(H_in, W_in):
    """Boxes crossing the border by more than 16 px, zero-area, inverted, sub-pixel, far outside."""
    return np.array([
        [-40.0, -40.0, 60.0, 60.0], [W_in - 50.0, H_in - 30.0, W_in + 45.0, H_in + 70.0],
        [100.0, 100.0, 100.0, 100.0], [200.0, 220.0, 200.0, 300.0],
        [300.0, 300.0, 250.0, 240.0], [17.3, 19.9, 17.9, 20.4],
        [-500.0, -500.0, -300.0, -350.0], [W_in + 100.0, 5.0, W_in + 200.0, 90.0],
        [0.0, 0.0, float(W_in), float(H_in)], [64.0, 64.0, 96.0, 96.0],
        [-32.0, 32.0, 0.0, 64.0], [W_in - 32.0, H_in - 32.0, float(W_in), float(H_in)],
    ], dtype=np.float64)


class Scene:
    """A set of moving identities producing per-frame detections (embs, boxes, confs).

    ``churn`` > 0 replaces that fraction of identities every ``churn_every`` frames, and
    ``drop`` hides each identity with that probability per frame, so births, misses,
    re-activation, long-lost ReID and purges are all exercised.
    """

    def __init__(self, seed, n, H_in, W_in, drop=0.0, churn=0.0, churn_every=20,
                 noise=0.05, shuffle=True):
        self.rng = np.random.default_rng(seed)
        self.n, self.H, self.W = n, H_in, W_in
        self.drop, self.churn, self.churn_every = drop, churn, churn_every
        self.noise, self.shuffle = noise, shuffle
        self.base = self.rng.standard_normal((n, 128))
        self.boxes = random_boxes(self.rng, n, H_in, W_in)
        self.frame = 0

    def _move(self):
        r = self.rng
        cx = (self.boxes[:, 0] + self.boxes[:, 2]) / 2 + r.normal(0, 2.0, self.n)
        cy = (self.boxes[:, 1] + self.boxes[:, 3]) / 2 + r.normal(0, 2.0, self.n)
        w = np.maximum(self.boxes[:, 2] - self.boxes[:, 0] + r.normal(0, 0.4, self.n), 8.0)
        h = np.maximum(self.boxes[:, 3] - self.boxes[:, 1] + r.normal(0, 0.4, self.n), 8.0)
        self.boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], axis=1)

    def step(self):
        """Returns the reference's ``obj`` dict (tracking.py:317-323) for the next frame."""
        r = self.rng
        if self.frame > 0:
            self._move()
            if self.churn > 0 and self.frame % self.churn_every == 0:
                k = max(1, int(self.churn * self.n))
                idx = r.choice(self.n, k, replace=False)
                self.base[idx] = r.standard_normal((k, 128))
                self.boxes[idx] = random_boxes(r, k, self.H, self.W)
        keep = np.nonzero(r.uniform(size=self.n) >= self.drop)[0]
        if self.shuffle:
            keep = r.permutation(keep)
        e = self.base[keep] + self.noise * r.standard_normal((len(keep), 128))
        e = (e / np.linalg.norm(e, axis=1, keepdims=True)).astype(np.float32)
        confs = r.uniform(0.56, 0.99, len(keep))
        obj = {
            "embs": [e[i] for i in range(len(keep))],
            "bboxes": [[float(v) for v in self.boxes[k]] for k in keep],
            "confs": [float(v) for v in confs],
            "input_hw": (self.H, self.W),
            "frame_id": self.frame,
        }
        self.frame += 1
        return obj


def lsap_matrix(rng, m, n, gated=0.0):
    """U[0,2) float32 matrix (unique optimum w.p. 1); ``gated`` fraction set to 1e9."""
    c = rng.uniform(0.0, 2.0, (m, n)).astype(np.float32)
    if gated > 0:
        c[rng.uniform(size=(m, n)) < gated] = np.float32(1e9)
    return c
