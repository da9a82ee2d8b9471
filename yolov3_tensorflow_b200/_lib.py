"""ctypes binding of libyolob200.so (the C ABI declared in include/yolob200.h).

The library is the product: there is no Python/PyTorch fallback.  Importing this
module when the shared object is missing raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libyolob200.so")

YB_F16, YB_BF16, YB_F32 = 0, 1, 2
YB_W_HWIO, YB_W_OIHW, YB_W_OHWI = 0, 1, 2
YB_OPT_SGD, YB_OPT_MOMENTUM, YB_OPT_RMSPROP, YB_OPT_ADAM = 0, 1, 2, 3
YB_TRAIN_FORWARD_ONLY, YB_TRAIN_BN_FROZEN, YB_TRAIN_NO_BACKWARD = 1, 2, 4


class YoloB200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with ./build.sh (or python -c 'import __graft_entry__ as g; g.build()'). "
        "yolov3_tensorflow_b200 has no CPU/PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class ConvDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("n", "h", "w", "cin", "cout", "ksize", "stride", "in_ld", "out_ld", "res_ld",
                                   "dtype", "out_fp32", "leaky", "upsample2x")]


class LayerInfo(C.Structure):
    _fields_ = [(n, i32) for n in ("index", "cin", "cout", "ksize", "stride", "has_bn", "in_h", "in_w", "out_h",
                                   "out_w", "is_head", "scope_index", "upsample2x")]


class Optimizer(C.Structure):      # yb_optimizer
    _fields_ = [("kind", i32)] + [(n, f32) for n in ("lr", "grad_scale", "momentum", "decay", "beta1", "beta2", "epsilon",
                                                      "weight_decay", "clip_norm")]


_SIGS = {
    "yb_version": ([], i32),
    "yb_last_error_string": ([], C.c_char_p),
    "yb_set_option": ([C.c_char_p, C.c_char_p], i32),
    "yb_get_option": ([C.c_char_p], C.c_char_p),
    "yb_device_info": ([C.POINTER(i32)] * 3, i32),
    "yb_conv2d_fwd": ([C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp], i32),
    "yb_conv_cout_pad": ([i32], i32),
    "yb_debug_set_conv_trace": ([vp], i32),
    "yb_stem_conv_fwd": ([vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp], i32),
    "yb_conv3x3_thin_fwd": ([C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp], i32),
    "yb_conv3x3_halo_supported": ([C.POINTER(ConvDesc)], i32),
    "yb_conv3x3_halo_fwd": ([C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp], i32),
    "yb_stem_conv1_fused_fwd": ([C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp], i32),
    "yb_stem_conv_fwd_tc": ([vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp], i32),
    "yb_stem_conv_fwd_tc_stats": ([vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "yb_process_box": ([vp, vp, vp, i32, i32, i32, i32, i32, C.POINTER(f32), vp, vp, vp, vp], i32),
    "yb_letterbox_params": ([i32, i32, i32, i32, C.POINTER(C.c_double), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], i32),
    "yb_letterbox_normalize": ([vp, i32, i32, C.c_long, i32, i32, vp, vp], i32),
    "yb_pack_conv_weights": ([vp, i32, i32, i32, i32, i32, i32, vp, vp], i32),
    "yb_bn_fold": ([vp, vp, vp, vp, i32, f32, vp, vp, vp], i32),
    "yb_conv2d_wgrad": ([C.POINTER(ConvDesc), vp, vp, i32, i32, vp, vp], i32),
    "yb_stem_conv_wgrad": ([vp, vp, i32, i32, i32, i32, vp, vp], i32),
    "yb_stem_conv_wgrad_tc": ([vp, vp, i32, i32, i32, i32, vp, vp], i32),
    "yb_pack_dgrad_weights": ([vp, i32, i32, i32, i32, i32, i32, vp, vp], i32),
    "yb_pack_dgrad_weights_s2": ([vp, i32, i32, i32, i32, i32, vp, vp], i32),
    "yb_conv2d_dgrad_s2": ([vp, vp, i32, i32, vp, vp, i32, vp, i32, vp], i32),
    "yb_bn_finalize": ([vp, vp, C.c_long, i32, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, vp], i32),
    "yb_bn_act_apply": ([vp, C.c_long, vp, vp, vp, C.c_long, vp, C.c_long, i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "yb_wgrad_split_plan": ([C.c_long, C.c_long, i32, i32, C.POINTER(C.c_long), C.POINTER(C.c_long)], i32),
    "yb_bn_stats_act_apply": ([vp, C.c_long, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, vp, C.c_long, vp, C.c_long,
                               i32, i32, i32, i32, i32, i32, i32, vp], i32),
    "yb_bn_bwd_reduce_workspace_bytes": ([C.POINTER(sz)], i32),
    "yb_bn_bwd_reduce": ([vp, C.c_long, vp, C.c_long, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp], i32),
    "yb_bn_bwd_apply": ([vp, C.c_long, vp, C.c_long, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, C.c_long, vp], i32),
    "yb_col_sum": ([vp, C.c_long, C.c_long, i32, i32, vp, vp], i32),
    "yb_col_stats": ([vp, C.c_long, C.c_long, i32, i32, vp, vp, vp], i32),
    "yb_reorg_layer": ([vp, i32, i32, i32, i32, i32, i32, C.POINTER(f32), vp, vp, vp, vp, vp], i32),
    "yb_predict": ([vp, vp, vp, i32, i32, i32, i32, C.POINTER(f32), vp, vp, vp, vp, vp], i32),
    "yb_nms_workspace_bytes": ([i32, i32, i32, i32, C.POINTER(sz)], i32),
    "yb_nms": ([vp, vp, i32, i32, i32, i32, f32, f32, vp, sz, vp, vp, vp, vp, vp, vp], i32),
    "yb_loss_workspace_bytes": ([i32, i32, i32, C.POINTER(sz)], i32),
    "yb_loss_layer": ([vp, vp, i32, i32, i32, i32, i32, i32, C.POINTER(f32), i32, i32, f32, f32, vp, sz, vp, vp, i32, i32, vp], i32),
    "yb_loss_finalize": ([vp, vp, vp], i32),
    "yb_box_iou": ([vp, vp, C.c_long, i32, vp, vp], i32),
    "yb_net_create": ([C.POINTER(vp), i32, i32, i32, i32, i32, i32], i32),
    "yb_net_destroy": ([vp], i32),
    "yb_net_num_layers": ([vp], i32),
    "yb_net_layer_info": ([vp, i32, C.POINTER(LayerInfo)], i32),
    "yb_net_arena_bytes": ([vp, C.POINTER(sz), C.POINTER(sz)], i32),
    "yb_net_bind": ([vp, vp, sz, vp, sz, vp], i32),
    "yb_net_refold_bn": ([vp, vp], i32),
    "yb_net_set_conv_params": ([vp, i32, vp, i32, vp, vp, vp, vp, vp, vp], i32),
    "yb_net_forward": ([vp, vp, vp, vp, vp, vp], i32),
    "yb_net_detect_supported": ([vp], i32),
    "yb_net_detect_workspace_bytes": ([vp, i32, C.POINTER(sz)], i32),
    "yb_net_detect": ([vp, vp, C.POINTER(f32), i32, f32, f32, vp, sz, vp, vp, vp, vp, vp, vp, vp], i32),
    "yb_net_detect_phases": ([vp, vp, C.POINTER(f32), i32, f32, f32, vp, sz, vp, vp, vp, vp, vp, vp, i32, vp], i32),
    "yb_net_forward_layers": ([vp, vp, vp, vp, vp, i32, i32, vp], i32),
    "yb_net_train_fwd_bwd": ([vp, vp, vp, vp, vp, C.POINTER(f32), i32, i32, f32, f32, vp, vp, vp, vp, i32, vp], i32),
    "yb_net_train_backward": ([vp, vp, i32, i32, i32, vp], i32),
    "yb_net_grad_range": ([vp, i32, i32, C.POINTER(vp), C.POINTER(sz)], i32),
    "yb_net_grad_buffer": ([vp, C.POINTER(vp), C.POINTER(sz)], i32),
    "yb_net_train_update": ([vp, C.POINTER(Optimizer), vp], i32),
    "yb_net_train_reset_state": ([vp, i32, vp], i32),
    "yb_net_opt_state": ([vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(i32), C.POINTER(vp)], i32),
    "yb_net_set_trainable": ([vp, i32, i32, vp], i32),
    "yb_net_train_refresh_dgrad": ([vp, vp], i32),
    "yb_net_get_conv_params": ([vp, i32] + [C.POINTER(vp)] * 6, i32),
    "yb_net_layer_grad": ([vp, i32] + [C.POINTER(vp)] * 4, i32),
    "yb_net_train_buffer": ([vp, i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], i32),
    "yb_net_layer_output": ([vp, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32)], i32),
    "yb_net_forward_launches": ([vp], i32),
}
for _name, (_args, _ret) in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _ret

EXPORTED = tuple(_SIGS)


def check(rc: int, what: str = ""):
    """Translate a yb_status into the Python exceptions the reference API raises."""
    if rc == 0:
        return
    msg = lib.yb_last_error_string().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise YoloB200Error(f"{what}: status {rc}: {msg}")


def set_option(key: str, value):
    """Runtime switch of the library (include/yolob200.h: yb_set_option); value None restores the default."""
    check(lib.yb_set_option(key.encode(), None if value is None else str(value).encode()), f"yb_set_option({key})")


def get_option(key: str) -> str:
    return lib.yb_get_option(key.encode()).decode()


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_handle():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def fptr(values):
    arr = (f32 * len(values))(*[float(v) for v in values])
    return arr
