"""utils/data_aug.py of the reference, the one function on the inference input path: `letterbox_resize` (:274-293),
fused with the caller's BGR->RGB + float32 / 255 (test_single_image.py:44-46) into one device kernel
(libyolob200.so: yb_letterbox_normalize).  Nearest-neighbour like the reference's default interp=0, bit-exact vs
cv2.resize(..., interpolation=0); the random augmentations of training are CPU image I/O and out of scope."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .._lib import lib, check, ptr, stream_handle


def letterbox_params(ori_height, ori_width, new_width, new_height):
    """(resize_ratio, resize_w, resize_h, dw, dh) of letterbox_resize — what test_single_image.py:64-66 needs to map
    the detections back to the original image."""
    ratio = C.c_double()
    rh, rw, dh, dw = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib.yb_letterbox_params(int(ori_height), int(ori_width), int(new_height), int(new_width), C.byref(ratio),
                                  C.byref(rh), C.byref(rw), C.byref(dh), C.byref(dw)), "yb_letterbox_params")
    return ratio.value, rw.value, rh.value, dw.value, dh.value


def letterbox_preprocess(img_bgr, new_width, new_height, device=None, out=None):
    """img_bgr: uint8 [H, W, 3] in OpenCV's BGR order (numpy array or torch tensor, host or CUDA) ->
    (x float32 [1, new_height, new_width, 3] RGB in [0, 1] on the device, resize_ratio, dw, dh):
    letterbox_resize(img, new_width, new_height) + cvtColor(BGR2RGB) + np.float32 + / 255. of the reference."""
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    t = torch.from_numpy(np.ascontiguousarray(img_bgr)) if isinstance(img_bgr, np.ndarray) else img_bgr
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError(f"letterbox_preprocess expects a uint8 [H, W, 3] image, got {t.dtype} {tuple(t.shape)}")
    t = t.to(dev, non_blocking=True).contiguous()
    h, w = int(t.shape[0]), int(t.shape[1])
    ratio, rw, rh, dw, dh = letterbox_params(h, w, new_width, new_height)
    if out is None:
        out = torch.empty((1, int(new_height), int(new_width), 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.yb_letterbox_normalize(ptr(t), h, w, 3 * w, int(new_height), int(new_width), ptr(out), stream_handle()),
              "yb_letterbox_normalize")
    return out, ratio, dw, dh
