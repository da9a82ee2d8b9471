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


def gen_eval_case(seed, n, w, h, class_num, max_gt=6):
    """Synthetic evaluation batch: ground truth from the oracle's generator, predictions = background
    boxes with low confidence plus, per gt box, three jittered copies (two with the right class at different
    confidences, one with a wrong class) and one well-overlapping duplicate that NMS must remove.
    -> (y_pred = [boxes [n,B,4], confs [n,B,1], probs [n,B,C]] float32, y_true = [y13, y26, y52], gts [(boxes, labels)])."""
    from oracle import yolov3_oracle as O
    rng = np.random.default_rng(seed)
    B = 3 * sum((h // s) * (w // s) for s in (32, 16, 8))
    boxes = np.zeros((n, B, 4), np.float32); confs = np.zeros((n, B, 1), np.float32); probs = np.zeros((n, B, class_num), np.float32)
    ys, gts = [[], [], []], []
    for i in range(n):
        gb, gl = O.synth_gt(rng, w, h, class_num, max_gt)
        y = O.process_box(gb, gl, [w, h], class_num, O.COCO_ANCHORS)
        for j in range(3):
            ys[j].append(y[j])
        gts.append((gb, gl))
        cx, cy = rng.uniform(0, w, B), rng.uniform(0, h, B)
        bw, bh = rng.uniform(4, 40, B), rng.uniform(4, 40, B)
        boxes[i] = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1)
        confs[i, :, 0] = rng.uniform(0.0, 0.25, B)
        probs[i] = rng.uniform(0.0, 0.5, (B, class_num))
        slots = rng.permutation(B)
        k = 0
        for (x0, y0, x1, y1, _), lab in zip(gb, gl):
            for conf, cls, jit in ((0.95, lab, 1.0), (0.8, lab, 2.5), (0.9, (lab + 1) % class_num, 1.5), (0.7, lab, 0.5)):
                d = rng.uniform(-jit, jit, 4)
                boxes[i, slots[k]] = (x0 + d[0], y0 + d[1], x1 + d[2], y1 + d[3])
                confs[i, slots[k], 0] = conf
                probs[i, slots[k]] = 0.01
                probs[i, slots[k], cls] = 0.9
                k += 1
    return [boxes, confs, probs], [np.stack(y) for y in ys], gts
