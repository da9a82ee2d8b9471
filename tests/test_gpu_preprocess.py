"""GPU tests of the device-side pre-processing: yb_process_box and yb_letterbox_normalize, bit-exact
against the reference-generated golden vectors (tests/golden/make_golden_preprocess.py) and against the oracle on
random cases."""
import os

import numpy as np
import pytest
import torch

from oracle import yolov3_oracle as O

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


def test_process_box_matches_reference_golden_with_collisions(golden_dir):
    from yolov3_tensorflow_b200.utils import data_utils as D
    g = _golden(golden_dir)
    W, H, C = (int(v) for v in g["pb_shape"])
    hb, hl, hc = D.pack_gt([g[f"pb_boxes{i}"] for i in range(3)], [g[f"pb_labels{i}"] for i in range(3)])
    ys = D.process_box_batch(hb, hl, hc, [W, H], C, O.COCO_ANCHORS)
    for y, name in zip(ys, ("y13", "y26", "y52")):
        ref = np.zeros(tuple(g[f"pb_{name}_shape"]), np.float32); ref[..., -1] = 1.0
        ref.reshape(-1)[g[f"pb_{name}_idx"]] = g[f"pb_{name}_val"]
        assert np.array_equal(y.cpu().numpy(), ref), name
    # the reference's single-image signature
    y1 = D.process_box(g["pb_boxes1"], g["pb_labels1"], [W, H], C, O.COCO_ANCHORS)
    assert np.array_equal(y1[2].cpu().numpy(), ys[2][1].cpu().numpy())


@pytest.mark.parametrize("n,w,h,cn,vmax", [(1, 416, 416, 80, 50), (5, 608, 608, 80, 50), (3, 160, 96, 20, 7), (2, 64, 64, 3, 200)])
def test_process_box_matches_oracle_random(n, w, h, cn, vmax):
    """Random ground truth incl. empty images, crowded images (many collisions at 64x64 with 200 boxes), odd element
    counts (n = 1 @416: 169 * 258 floats is not a multiple of 4) — bit-exact vs oracle.process_box (pinned to the
    reference by tests/test_oracle_golden.py)."""
    from yolov3_tensorflow_b200.utils import data_utils as D
    rng = np.random.default_rng(100 + n + w)
    bl, ll = [], []
    for i in range(n):
        boxes, labels = O.synth_gt(rng, w, h, cn, vmax)
        boxes[:, 4] = rng.uniform(0.5, 1.0, len(boxes)).astype(np.float32)
        if i == 1:
            boxes, labels = boxes[:0], labels[:0]
        bl.append(boxes); ll.append(labels)
    hb, hl, hc = D.pack_gt(bl, ll, vmax)
    ys = D.process_box_batch(hb, hl, hc, [w, h], cn, O.COCO_ANCHORS)
    for i in range(n):
        ref = O.process_box(bl[i], ll[i], [w, h], cn, O.COCO_ANCHORS)
        for y, r in zip(ys, ref):
            assert np.array_equal(y[i].cpu().numpy(), r)


def test_process_box_feeds_the_loss():
    """The device-built y_true drives compute_loss to the same values as the host-built one."""
    import yolov3_tensorflow_b200 as pkg
    from yolov3_tensorflow_b200.utils import data_utils as D
    rng = np.random.default_rng(3)
    n, w, h, cn = 2, 96, 64, 80
    bl, ll = zip(*[O.synth_gt(rng, w, h, cn, 6) for _ in range(n)])
    ys = D.process_box_batch(*D.pack_gt(list(bl), list(ll)), [w, h], cn, O.COCO_ANCHORS)
    host = [np.stack([O.process_box(b, l, [w, h], cn, O.COCO_ANCHORS)[j] for b, l in zip(bl, ll)]) for j in range(3)]
    m = pkg.yolov3(cn, O.COCO_ANCHORS)
    m.init_params(1)
    fms = m.forward(torch.rand((n, h, w, 3), device="cuda"))
    a = [float(v) for v in m.compute_loss(fms, list(ys))]
    b = [float(v) for v in m.compute_loss(fms, [torch.from_numpy(t).cuda() for t in host])]
    assert a == b


def test_process_box_rejects_bad_arguments():
    from yolov3_tensorflow_b200.utils import data_utils as D
    hb, hl, hc = D.pack_gt([np.zeros((2, 5), np.float32)], [np.zeros(2, np.int64)])
    with pytest.raises(ValueError):
        D.process_box_batch(hb, hl, hc, [100, 96], 80, O.COCO_ANCHORS)      # width not a multiple of 32
    with pytest.raises(ValueError):
        D.pack_gt([np.zeros((2, 5), np.float32)], [np.zeros(3, np.int64)])


def test_letterbox_matches_reference_golden(golden_dir):
    from yolov3_tensorflow_b200.utils import data_aug as A
    g = _golden(golden_dir)
    for i, (sh, sw, nw, nh) in enumerate(g["letterbox_cases"]):
        x, ratio, dw, dh = A.letterbox_preprocess(g[f"lb_src{i}"], int(nw), int(nh))
        assert tuple(x.shape) == (1, int(nh), int(nw), 3)
        assert np.array_equal(x.cpu().numpy(), g[f"lb_out{i}"]), f"case {i}"
        assert (ratio, dw, dh) == tuple(g[f"lb_meta{i}"].tolist())


@pytest.mark.parametrize("sh,sw,nw,nh", [(480, 640, 416, 416), (1080, 1920, 608, 608), (333, 500, 416, 416), (500, 333, 416, 416)])
def test_letterbox_matches_oracle_at_baseline_sizes(sh, sw, nw, nh):
    from yolov3_tensorflow_b200.utils import data_aug as A
    img = np.random.default_rng(sh + sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    x, ratio, dw, dh = A.letterbox_preprocess(img, nw, nh)
    rx, rr, rdw, rdh = O.letterbox_preprocess(img, nw, nh)
    assert np.array_equal(x.cpu().numpy(), rx) and (ratio, dw, dh) == (rr, rdw, rdh)
