"""The hot-path subset of the reference's utils/misc_utils.py: anchor / class-name parsers and
the darknet `.weights` loader (utils/misc_utils.py:31-47, 70-126)."""
from __future__ import annotations

import os

import numpy as np


def parse_anchors(anchor_path):
    """utils/misc_utils.py:31-37 -> float32 [N,2] (w,h)."""
    with open(anchor_path, "r") as f:
        return np.reshape(np.asarray(f.readline().split(","), np.float32), [-1, 2])


def read_class_names(class_name_path):
    """utils/misc_utils.py:40-45 -> {id: name}."""
    names = {}
    with open(class_name_path, "r") as data:
        for i, name in enumerate(data):
            names[i] = name.strip("\n")
    return names


DEFAULT_ANCHOR_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "yolo_anchors.txt")
DEFAULT_CLASS_NAME_PATH = os.path.join(os.path.dirname(DEFAULT_ANCHOR_PATH), "coco.names")


def load_weights(model, weights_file):
    """utils/misc_utils.py:70-126.  The reference takes the TF variable list; here the first argument
    is the `yolov3` model, whose conv table *is* that creation order.  Stream: 5 x int32 header, then per
    conv [beta, gamma, mean, var] (or [bias]) followed by the weights as (Cout, Cin, kh, kw) float32.
    The (Cout,Cin,kh,kw) -> engine-layout transposition the reference does on the host (:117-120)
    happens on the GPU (yb_pack_conv_weights).  Raises ValueError on a size mismatch
    (tf.assign(validate_shape=True))."""
    with open(weights_file, "rb") as fp:
        np.fromfile(fp, dtype=np.int32, count=5)
        weights = np.fromfile(fp, dtype=np.float32)
    ptr = 0
    params = []
    for cin, cout, k, s, bn in model.conv_table(model.class_num):
        p = {}
        names = ("beta", "gamma", "mean", "var") if bn else ("b",)
        for name in names:
            if ptr + cout > weights.size:
                raise ValueError("darknet weights file is too short for this architecture")
            p[name] = weights[ptr:ptr + cout]
            ptr += cout
        n = cout * cin * k * k
        if ptr + n > weights.size:
            raise ValueError("darknet weights file is too short for this architecture")
        p["w"] = weights[ptr:ptr + n].reshape(cout, cin, k, k)
        ptr += n
        params.append(p)
    if ptr != weights.size:
        raise ValueError(f"darknet weights file has {weights.size} floats, architecture needs {ptr}")
    model.set_params(params, layout="OIHW")
    return ptr


def save_weights(params, weights_file, layout="HWIO"):
    """Inverse of load_weights: write 75 parameter dicts as a darknet stream."""
    with open(weights_file, "wb") as f:
        np.array([0, 2, 0, 0, 0], np.int32).tofile(f)
        for p in params:
            if "gamma" in p:
                for k in ("beta", "gamma", "mean", "var"):
                    np.asarray(p[k], np.float32).tofile(f)
            else:
                np.asarray(p["b"], np.float32).tofile(f)
            w = np.asarray(p["w"], np.float32)
            if layout == "HWIO":
                w = np.transpose(w, (3, 2, 0, 1))
            np.ascontiguousarray(w).tofile(f)
