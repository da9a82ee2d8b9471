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


# ------------------------------------------------------------------------- N1: LR schedules / optimizer selection
def _lr_args(**kw):
    import types
    base = dict(lr_type="exponential", learning_rate_init=1e-3, lr_decay_freq=400, lr_decay_factor=0.96, lr_lower_bound=1e-6,
                total_epoches=10, use_warm_up=True, warm_up_epoch=3, train_batch_num=100,
                pw_boundaries=[300.0, 500.0], pw_values=[1e-3, 3e-4, 1e-4])
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_learning_rate_schedules_match_the_oracle_restatement():
    """utils/misc_utils.py:129-148 + warm-up train.py:93-99: the host-side schedule functions against the oracle's
    independent restatement (oracle.learning_rate) over all five lr types, with and without warm-up."""
    from oracle import yolov3_oracle as O
    from yolov3_tensorflow_b200.utils import misc_utils as M
    steps = [0, 1, 150, 299, 300, 301, 399, 400, 401, 799, 800, 1200, 2799, 2800, 5000]
    for kind in ("exponential", "cosine_decay", "cosine_decay_restart", "fixed", "piecewise"):
        for warm in (True, False):
            a = _lr_args(lr_type=kind, use_warm_up=warm)
            for s in steps:
                got, ref = M.learning_rate_at(a, s), O.learning_rate(a, s)
                assert abs(got - ref) <= 1e-12 + 1e-9 * abs(ref), (kind, warm, s, got, ref)
    assert M.learning_rate_at(_lr_args(), 150) == 1e-3 * 150 / 300                      # linear warm-up (train.py:95)
    assert M.config_learning_rate(_lr_args(lr_type="exponential"), 800) == max(1e-3 * 0.96 ** 2, 1e-6)
    assert M.config_learning_rate(_lr_args(lr_type="piecewise"), 300) == 1e-3 and M.config_learning_rate(_lr_args(lr_type="piecewise"), 300.5) == 3e-4
    with pytest.raises(ValueError, match="Unsupported learning rate type"):
        M.config_learning_rate(_lr_args(lr_type="poly"), 0)


def test_config_optimizer_and_tf_variable_names():
    from yolov3_tensorflow_b200.utils import misc_utils as M
    for name in ("momentum", "rmsprop", "adam", "sgd"):
        assert M.config_optimizer(name, 1e-3).name == name
    with pytest.raises(ValueError, match="Unsupported optimizer type"):
        M.config_optimizer("adagrad", 1e-3)                                                # utils/misc_utils.py:161
    names = M.tf_variable_names(80)
    assert len(names) == 72 * 5 + 3 * 2 == 366
    assert names[0][2] == "yolov3/darknet53_body/Conv/weights:0" and names[-1][2] == "yolov3/yolov3_head/Conv_22/biases:0"
    assert names[5][2] == "yolov3/darknet53_body/Conv_1/weights:0" and names[1][2] == "yolov3/darknet53_body/Conv/BatchNorm/gamma:0"


def test_wgrad_split_plan_one_wave_and_every_block_once():
    """Host logic of yb_conv2d_wgrad (csrc/conv_wgrad.cu: wgrad_pick_splits): the split-K count minimises
    waves x (pixel blocks per CTA + epilogue); every 64-pixel block is covered exactly once, no split is empty."""
    import ctypes as C
    from yolov3_tensorflow_b200 import _lib as L

    def plan(num_kb, tiles, sms=148, epi=40):
        s, k = C.c_long(), C.c_long()
        L.check(L.lib.yb_wgrad_split_plan(num_kb, tiles, sms, epi, C.byref(s), C.byref(k)), "yb_wgrad_split_plan")
        return s.value, k.value

    # the training step's layers at batch 32 @416 (pixel blocks, tiles) -> (splits, blocks per split)
    cases = {(21632, 1): (148, 147), (5408, 3): (49, 111), (1352, 6): (24, 57), (338, 24): (6, 57), (85, 96): (1, 85),
             (85, 32): (4, 22), (1352, 2): (72, 19)}
    for (num_kb, tiles), want in cases.items():
        assert plan(num_kb, tiles) == want, (num_kb, tiles, plan(num_kb, tiles))
    for num_kb, tiles in list(cases) + [(1, 1), (7, 400), (1000, 149), (64, 1)]:
        s, k = plan(num_kb, tiles)
        assert s >= 1 and (s - 1) * k < num_kb <= s * k
        if tiles <= 148 and num_kb >= 148 // tiles:
            assert tiles * s <= 148                 # one wave of one-CTA-per-SM blocks whenever the work allows it
    # ignoring the epilogue cost (the first version's behaviour) splits further
    assert plan(338, 24, epi=0)[0] > plan(338, 24)[0]
    with pytest.raises(Exception):
        plan(0, 1)
