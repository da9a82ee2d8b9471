"""CPU tests of the host-side mirror of the reference's utility functions (no GPU work): the numpy NMS variants
against the golden vectors produced by the reference's own code, anchor / class-name parsing, batch sharding and the
darknet weight file writer."""
import os

import numpy as np
import pytest

from oracle import yolov3_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pkg():
    import yolov3_tensorflow_b200 as pkg
    return pkg


def test_cpu_and_py_nms_match_reference_golden(golden_dir):
    pkg = _pkg()
    g = np.load(os.path.join(golden_dir, "nms.npz"))
    cb, cs, cl = pkg.cpu_nms(g["boxes_in"][None], g["scores_in"][None], 6, max_boxes=20, score_thresh=0.3, iou_thresh=0.45)
    assert np.array_equal(cb, g["cpu_boxes"]) and np.array_equal(cs, g["cpu_scores"]) and np.array_equal(cl, g["cpu_labels"])
    assert list(pkg.py_nms(g["boxes_in"], g["scores_in"][:, 0], max_boxes=30, iou_thresh=0.5)) == list(g["py_keep"])


def test_cpu_nms_returns_none_triplet_without_candidates():
    pkg = _pkg()
    boxes = np.zeros((1, 5, 4), np.float32)
    scores = np.zeros((1, 5, 3), np.float32)
    assert pkg.cpu_nms(boxes, scores, 3, score_thresh=0.5) == (None, None, None)   # utils/nms_utils.py:118-119


def test_anchor_and_class_name_files():
    pkg = _pkg()
    data = os.path.join(ROOT, "yolov3_tensorflow_b200", "data")
    anchors = pkg.parse_anchors(os.path.join(data, "yolo_anchors.txt"))
    assert anchors.shape == (9, 2) and anchors.dtype == np.float32
    assert np.array_equal(anchors, np.asarray(O.COCO_ANCHORS, np.float32))
    names = pkg.read_class_names(os.path.join(data, "coco.names"))
    assert len(names) == 80 and names[0] == "person" and names[79] == "toothbrush"


def test_shard_batch_covers_the_batch_once():
    from yolov3_tensorflow_b200.parallel import shard_batch
    for n, world in ((256, 8), (64, 2), (64, 1)):
        spans = [shard_batch(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert len({hi - lo for lo, hi in spans}) == 1
    with pytest.raises(ValueError):          # equal shards only: every loss term is a mean over the local batch
        shard_batch(10, 0, 4)


def test_save_weights_writes_the_darknet_stream(tmp_path):
    from yolov3_tensorflow_b200.utils.misc_utils import save_weights
    params = O.make_params(80, seed=11, random_bn=True)
    path = tmp_path / "w.weights"
    save_weights(params, str(path), layout="HWIO")
    assert os.path.getsize(path) == 20 + 4 * 62_001_757
    back = O.load_darknet_weights(str(path), 80)          # the oracle's reader (utils/misc_utils.py:70-126 order)
    for a, b in zip(params, back):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
