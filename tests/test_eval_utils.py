"""Evaluation callers of gpu_nms (utils/eval_utils.py of the reference) against golden vectors
produced by the reference's own evaluate_on_gpu / get_preds_gpu / voc_eval (tests/golden/make_golden_eval.py).
CPU tests replace the device NMS by the oracle's; the GPU tests run the real batched NMS."""
import os

import numpy as np
import pytest

from oracle import yolov3_oracle as O
from tests.synth import gen_eval_case

NMS = dict(max_boxes=20, score_thresh=0.3, nms_thresh=0.45)


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "eval.npz"))


def _oracle_nms_batch(y_pred, num_classes, max_boxes, score_thresh, nms_thresh):
    out = []
    for i in range(y_pred[0].shape[0]):
        b, s, l, _ = O.gpu_nms(y_pred[0][i:i + 1], (y_pred[1][i:i + 1] * y_pred[2][i:i + 1]).astype(np.float32), num_classes,
                               max_boxes, score_thresh, nms_thresh)
        out.append((b, s, l))
    return out


def _check_eval(E, g, tag):
    seed, n, w, h, cn = (int(v) for v in g[f"ev_{tag}_cfg"])
    y_pred, y_true, gts = gen_eval_case(seed, n, w, h, cn)
    tp, tr, pr = E.evaluate_on_gpu(y_pred, y_true, cn, 0.5, calc_now=False, **NMS)
    assert [tp[i] for i in range(cn)] == g[f"ev_{tag}_tp"].tolist()
    assert [tr[i] for i in range(cn)] == g[f"ev_{tag}_true"].tolist()
    assert [pr[i] for i in range(cn)] == g[f"ev_{tag}_pred"].tolist()
    assert sum(tp.values()) > 0 and sum(tp.values()) < sum(pr.values())          # the case has hits, misses and wrong classes
    rec, prec = E.evaluate_on_gpu(y_pred, y_true, cn, 0.5, calc_now=True, **NMS)
    assert (rec, prec) == tuple(g[f"ev_{tag}_rp"].tolist())
    preds = E.get_preds_gpu([100 + i for i in range(n)], y_pred, cn, **NMS)
    got = np.asarray([[float(v) for v in row] for row in preds], np.float64).reshape(-1, 7)
    assert np.array_equal(got, g[f"pr_{tag}"])
    gt_dict = {100 + i: [[float(v) for v in b[:4]] + [int(l)] for b, l in zip(*gts[i])] for i in range(n)}
    for row in g[f"voc_{tag}"]:
        c, m07 = int(row[0]), bool(row[1])
        r = E.voc_eval(gt_dict, preds, c, iou_thres=0.5, use_07_metric=m07)
        assert np.array_equal(np.asarray([float(v) for v in r]), row[2:], equal_nan=True), (c, m07, r, row[2:])   # (npos = 0 -> nan, like the reference)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_eval_host_logic_matches_reference(golden_dir, tag, monkeypatch):
    from yolov3_tensorflow_b200.utils import eval_utils as E
    monkeypatch.setattr(E, "_nms_batch", _oracle_nms_batch)
    _check_eval(E, _golden(golden_dir), tag)


def test_calc_iou_and_voc_ap_known_answers():
    from yolov3_tensorflow_b200.utils import eval_utils as E
    iou = E.calc_iou(np.array([[0., 0., 2., 2.]]), np.array([[1., 1., 3., 3.], [4., 4., 5., 5.]]))
    assert abs(iou[0, 0] - 1.0 / 7.0) < 1e-9 and iou[0, 1] == 0.0
    rec, prec = np.array([0.25, 0.5, 0.5, 1.0]), np.array([1.0, 1.0, 2.0 / 3.0, 0.8])
    assert abs(E.voc_ap(rec, prec) - (0.25 * 1.0 + 0.25 * 1.0 + 0.5 * 0.8)) < 1e-12
    assert abs(E.voc_ap(rec, prec, True) - (6 * 1.0 + 5 * 0.8) / 11.0) < 1e-12


def test_parse_gt_rec_lines_letterbox():
    from yolov3_tensorflow_b200.utils import eval_utils as E
    gt = E.parse_gt_rec_lines([(7, [[10., 20., 110., 220.]], [3], 500, 375)], (416, 416), True)
    r = min(416 / 500, 416 / 375); dw = int((416 - int(r * 500)) / 2); dh = int((416 - int(r * 375)) / 2)
    assert gt[7] == [[10. * r + dw, 20. * r + dh, 110. * r + dw, 220. * r + dh, 3]]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_eval_on_gpu_matches_reference(golden_dir, tag):
    """The same golden through the real device path: ONE batched gpu_nms call for the whole batch."""
    import torch
    from yolov3_tensorflow_b200.utils import eval_utils as E
    assert torch.cuda.is_available()
    _check_eval(E, _golden(golden_dir), tag)
