"""Seeded synthetic-input generators shared by tests/golden/make_golden.py, the tests
and bench.py (numpy default_rng / PCG64 streams are stable across numpy versions)."""
import numpy as np


def gen_inputs(seed, n, h, w):
    """Images U[0,1) float32 NHWC."""
    return np.random.default_rng(seed).random((n, h, w, 3), dtype=np.float32)


def gen_fms(seed, n, h, w, class_num, scale=1.5):
    """Random detection-head logits for the three scales (/32, /16, /8)."""
    rng = np.random.default_rng(seed)
    D = 3 * (5 + class_num)
    return [(rng.standard_normal((n, h // s, w // s, D)) * scale).astype(np.float32) for s in (32, 16, 8)]


def gen_nms_boxes(seed, B, class_num, extent=416.0, dense=False, lo=4.0, hi=200.0):
    """SURVEY.md §8d cfg 5: boxes xyxy float32 [B,4], scores [B,C] (sparse u1*u2 or dense U[0,1))."""
    rng = np.random.default_rng(seed)
    cx, cy = rng.uniform(0, extent, B), rng.uniform(0, extent, B)
    bw = np.exp(rng.uniform(np.log(lo), np.log(hi), B))
    bh = np.exp(rng.uniform(np.log(lo), np.log(hi), B))
    boxes = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1).astype(np.float32)
    if dense:
        scores = rng.random((B, class_num), dtype=np.float32)
    else:
        scores = (rng.random((B, class_num), dtype=np.float32) * rng.random((B, class_num), dtype=np.float32))
    return boxes, scores
