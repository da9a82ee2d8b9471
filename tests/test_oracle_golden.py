"""CPU tests: pin oracle/ against the golden vectors that the reference's own Python
sources produced (tests/golden/make_golden.py), and against the self-made KATs of
SURVEY.md §8c (parameter count, FLOP count)."""
import os

import numpy as np
import pytest
import torch

from oracle import yolov3_oracle as O
from tests.synth import gen_fms, gen_inputs


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_kat_param_and_flop_counts():
    assert len(O.conv_specs()) == 75
    assert O.count_params(80) == 62_001_757           # payload of the public yolov3.weights
    assert abs(O.forward_flops(416, 416) / 1e9 - 65.864) < 1e-3   # darknet prints 65.86 BFLOPs
    assert abs(O.forward_flops(608, 608) / 1e9 - 140.692) < 1e-3
    trainable = sum(p.size for q in O.make_params(80, 0) for k, p in q.items() if k in ("w", "gamma", "beta", "b"))
    assert trainable == 61_949_149


def test_forward_inference_matches_reference_wiring(golden_dir):
    g = _load(golden_dir, "forward_infer.npz")
    n, h, w = g["shape"]
    params = O.make_params(80, seed=int(g["seed_params"]), random_bn=True)
    x = gen_inputs(int(g["seed_x"]), n, h, w)
    fms = O.forward(x, params)
    for a, name in zip(fms, ("fm1", "fm2", "fm3")):
        np.testing.assert_allclose(a, g[name], rtol=2e-4, atol=2e-5)
    b, c, p = O.predict(fms, O.COCO_ANCHORS, (h, w))
    np.testing.assert_allclose(b, g["boxes"], rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(c, g["confs"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(p, g["probs"], rtol=2e-4, atol=1e-6)


def test_forward_training_matches_reference_wiring(golden_dir):
    g = _load(golden_dir, "forward_train.npz")
    n, h, w = g["shape"]
    params = O.make_params(80, seed=int(g["seed_params"]), random_bn=True)
    x = gen_inputs(int(g["seed_x"]), n, h, w)
    fms, stats = O.forward(x, params, is_training=True, bn_decay=float(g["decay"]))
    assert len(stats) == int(g["n_stats"]) == 72
    for a, name in zip(fms, ("fm1", "fm2", "fm3")):
        np.testing.assert_allclose(a, g[name], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(stats[0][0].numpy(), g["mean_first"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(stats[0][1].numpy(), g["var_first"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(stats[-1][0].numpy(), g["mean_last"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(stats[-1][1].numpy(), g["var_last"], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("tag", ["c80", "c20"])
def test_decode_bit_exact(golden_dir, tag):
    g = _load(golden_dir, f"decode_{tag}.npz")
    n, h, w = g["shape"]
    cn = int(g["class_num"])
    f = gen_fms(int(g["seed"]), n, h, w, cn)
    b, c, p = O.predict(f, O.COCO_ANCHORS, (h, w), cn)
    # same numpy float32 op order as the reference graph -> bit exact
    assert np.array_equal(b, g["boxes"])
    assert np.array_equal(c, g["confs"])
    assert np.array_equal(p, g["probs"])
    xy, bx, _, _ = O.reorg_layer(f[0], O.COCO_ANCHORS[6:9], (h, w), cn)
    assert np.array_equal(xy, g["xy_offset"])
    assert np.array_equal(bx, g["reorg_boxes"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_loss_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"loss_{tag}.npz")
    n, h, w = g["shape"]
    cn = int(g["class_num"])
    f = gen_fms(int(g["seed_fm"]), n, h, w, cn, scale=1.0)
    y_true = [g["y_true_13"], g["y_true_26"], g["y_true_52"]]
    for ls in (False, True):
        for fo in (False, True):
            got = O.compute_loss([torch.from_numpy(a) for a in f], y_true, O.COCO_ANCHORS, (h, w), cn, ls, fo)
            got = np.array([float(v) for v in got])
            np.testing.assert_allclose(got, g[f"loss_ls{int(ls)}_fo{int(fo)}"], rtol=2e-5, atol=1e-6)
    # box_iou table
    anchors = O.COCO_ANCHORS[3:6]
    _, pb, _, _ = O.reorg_layer(f[1], anchors, (h, w), cn)
    yt = y_true[1][n - 1]
    valid = yt[..., 0:4][yt[..., 4] > 0]
    iou = O.box_iou(torch.from_numpy(pb[n - 1]), torch.from_numpy(valid)).numpy()
    np.testing.assert_allclose(iou, g["iou_scale2_lastimg"], rtol=1e-5, atol=1e-7)


def test_loss_gradient_matches_fp64_finite_differences():
    rng = np.random.default_rng(5)
    n, h, w, cn = 2, 64, 64, 4
    f = gen_fms(7, n, h, w, cn, scale=1.0)
    ys = [[], [], []]
    for i in range(n):
        boxes, labels = O.synth_gt(rng, w, h, cn, 6)
        y = O.process_box(boxes, labels, [w, h], cn, O.COCO_ANCHORS)
        for j in range(3):
            ys[j].append(y[j])
    y_true = [np.stack(y) for y in ys]
    for ls, fo in ((False, False), (True, True)):
        losses, grads = O.loss_and_grad(f, y_true, O.COCO_ANCHORS, (h, w), cn, ls, fo, dtype=torch.float64)

        def total(fs):
            return float(O.compute_loss([torch.tensor(a, dtype=torch.float64) for a in fs], y_true,
                                        O.COCO_ANCHORS, (h, w), cn, ls, fo)[0])
        # probe positive cells and random cells
        for s in range(3):
            pos = np.argwhere(y_true[s][..., 4] > 0)
            probes = [tuple(p) for p in pos[:3]]
            for nb, gy, gx, a in probes:
                for j in (0, 2, 4, 5 + int(np.argmax(y_true[s][nb, gy, gx, a, 5:-1]))):
                    ch = a * (5 + cn) + j
                    eps = 1e-6
                    fp = [a_.astype(np.float64).copy() for a_ in f]
                    fm_ = [a_.astype(np.float64).copy() for a_ in f]
                    fp[s][nb, gy, gx, ch] += eps
                    fm_[s][nb, gy, gx, ch] -= eps
                    num = (total(fp) - total(fm_)) / (2 * eps)
                    assert abs(num - grads[s][nb, gy, gx, ch]) < 1e-5 * max(1.0, abs(num)), (s, j, num)


def test_process_box_matches_reference(golden_dir):
    g = _load(golden_dir, "process_box.npz")
    n, h, w = g["shape"]
    for i in range(2):
        y = O.process_box(g[f"boxes{i}"], g[f"labels{i}"], [w, h], 80, O.COCO_ANCHORS)
        assert np.array_equal(y[0], g["y13"][i])
        assert np.array_equal(y[1], g["y26"][i])
        assert np.array_equal(y[2], g["y52"][i])


def test_gpu_nms_python_layer_matches_reference(golden_dir):
    g = _load(golden_dir, "nms.npz")
    b, s, l, idx = O.gpu_nms(g["boxes_in"][None], g["scores_in"][None], 6, 20, 0.3, 0.45)
    assert np.array_equal(b, g["gpu_boxes"])
    assert np.array_equal(s, g["gpu_scores"])
    assert np.array_equal(l, g["gpu_labels"])
    # orig indices really address the inputs
    assert np.array_equal(g["boxes_in"][idx], b)
    assert np.array_equal(g["scores_in"][idx, l], s)
    # the tie block (boxes 10..19 share a class-2 score): lower index wins
    t = idx[l == 2]
    tie = [int(i) for i in t if 10 <= i < 20]
    assert tie == sorted(tie)


def test_tf_nms_semantics_small_cases():
    # strict '>' : IoU == thr does not suppress
    boxes = np.array([[0, 0, 2, 2], [1, 0, 3, 2], [10, 10, 11, 11], [0, 0, 0, 5]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.95], np.float32)
    iou01 = 2.0 / 6.0
    assert list(O.tf_nms_cpu(boxes, scores, 10, np.float32(iou01))) == [3, 0, 1, 2]
    assert list(O.tf_nms_cpu(boxes, scores, 10, 0.3)) == [3, 0, 2]
    assert list(O.tf_nms_cpu(boxes, scores, 2, 0.3)) == [3, 0]
    assert list(O.tf_nms_cpu(boxes[:0], scores[:0], 5, 0.5)) == []
    # flipped corners are legal
    fl = np.array([[2, 2, 0, 0], [0, 0, 2, 2]], np.float32)
    assert list(O.tf_nms_cpu(fl, np.array([0.5, 0.6], np.float32), 5, 0.5)) == [1]


def test_darknet_weights_roundtrip(tmp_path):
    params = O.make_params(80, seed=3, random_bn=True)
    path = tmp_path / "w.weights"
    O.write_darknet_weights(str(path), params)
    assert os.path.getsize(path) == 20 + 4 * 62_001_757
    back = O.load_darknet_weights(str(path), 80)
    for a, b in zip(params, back):
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_c_nms_oracle_equals_numpy_oracle(golden_dir):
    from tests.synth import gen_nms_boxes
    g = _load(golden_dir, "nms.npz")
    for boxes, scores, cn, mb in [(g["boxes_in"], g["scores_in"], 6, 20)] + [
            gen_nms_boxes(5, 1500, 9, dense=d, extent=160.0) + (9, 40) for d in (False, True)]:
        a = O.gpu_nms(boxes[None], scores[None], cn, mb, 0.3, 0.45)
        b = O.gpu_nms_c(boxes[None], scores[None], cn, mb, 0.3, 0.45)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    b = O.gpu_nms_c(g["boxes_in"][None], g["scores_in"][None], 6, 20, 0.3, 0.45)
    assert np.array_equal(b[0], g["gpu_boxes"]) and np.array_equal(b[2], g["gpu_labels"])


def test_letterbox_preprocess_matches_reference(golden_dir):
    """oracle.letterbox_preprocess (numpy restatement of cv2's nearest-neighbour resize) vs the reference's own
    letterbox_resize + BGR2RGB + /255 run over OpenCV (tests/golden/make_golden_preprocess.py): bit-exact."""
    g = _load(golden_dir, "preprocess.npz")
    for i, (sh, sw, nw, nh) in enumerate(g["letterbox_cases"]):
        x, ratio, dw, dh = O.letterbox_preprocess(g[f"lb_src{i}"], int(nw), int(nh))
        assert np.array_equal(x, g[f"lb_out{i}"])
        assert (ratio, dw, dh) == tuple(g[f"lb_meta{i}"].tolist())


def test_process_box_collisions_match_reference(golden_dir):
    g = _load(golden_dir, "preprocess.npz")
    W, H, C = (int(v) for v in g["pb_shape"])
    ys = [[], [], []]
    for i in range(3):
        y = O.process_box(g[f"pb_boxes{i}"], g[f"pb_labels{i}"], [W, H], C, O.COCO_ANCHORS)
        for j in range(3):
            ys[j].append(y[j])
    for j, name in enumerate(("y13", "y26", "y52")):
        y = np.stack(ys[j], 0)
        ref = np.zeros(tuple(g[f"pb_{name}_shape"]), np.float32); ref[..., -1] = 1.0
        ref.reshape(-1)[g[f"pb_{name}_idx"]] = g[f"pb_{name}_val"]
        assert np.array_equal(y, ref)
