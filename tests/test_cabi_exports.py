"""CPU tests: the C-ABI library loads without a GPU and exports every symbol that
include/yolob200.h declares; the ctypes table in _lib.py covers the same set; argument
validation that needs no device work returns the documented status codes."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "yolob200.h")
LIB = os.path.join(ROOT, "yolov3_tensorflow_b200", "libyolob200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(LIB)


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in yolob200.h but not exported: {missing}"


def test_ctypes_table_matches_header():
    from yolov3_tensorflow_b200 import _lib
    assert sorted(_lib.EXPORTED) == _declared()


def test_version_and_error_string(lib):
    lib.yb_last_error_string.restype = C.c_char_p
    assert lib.yb_version() >= 100
    # argument validation happens before any CUDA call
    n = C.c_size_t()
    assert lib.yb_nms_workspace_bytes(1, 10, 0, 5, C.byref(n)) == -1
    assert b"nms" in lib.yb_last_error_string()
    assert lib.yb_nms_workspace_bytes(2, 100, 80, 200, C.byref(n)) == 0 and n.value > 2 * 100 * 80 * 8
    assert lib.yb_conv_cout_pad(255) == 256 and lib.yb_conv_cout_pad(64) == 64 and lib.yb_conv_cout_pad(1024) == 1024


def test_net_plan_is_host_only_until_bound(lib):
    """Creating a plan and querying its schedule needs no device: 75 layers in darknet order."""
    from yolov3_tensorflow_b200 import _lib
    from yolov3_tensorflow_b200.model import yolov3
    h = C.c_void_p()
    assert _lib.lib.yb_net_create(C.byref(h), 80, 2, 416, 416, 0, 0) == 0
    assert _lib.lib.yb_net_num_layers(h) == 75
    table = yolov3.conv_table(80)
    flops = 0
    for i in range(75):
        info = _lib.LayerInfo()
        assert _lib.lib.yb_net_layer_info(h, i, C.byref(info)) == 0
        assert (info.cin, info.cout, info.ksize, info.stride, bool(info.has_bn)) == table[i]
        flops += 2 * info.out_h * info.out_w * info.cout * info.cin * info.ksize ** 2
    assert abs(flops / 1e9 - 65.864) < 1e-3          # SURVEY.md §8c KAT
    a, p = C.c_size_t(), C.c_size_t()
    assert _lib.lib.yb_net_arena_bytes(h, C.byref(a), C.byref(p)) == 0
    assert p.value > 62_001_757 * 6                  # fp32 master + 16-bit packed copies
    assert _lib.lib.yb_net_create(C.byref(C.c_void_p()), 80, 1, 100, 416, 0, 0) == -1
    assert _lib.lib.yb_net_destroy(h) == 0


def test_conv_table_equals_oracle_walk():
    from oracle import yolov3_oracle as O
    from yolov3_tensorflow_b200.model import yolov3
    for cn in (80, 20, 1):
        assert yolov3.conv_table(cn) == [(a[1], a[2], a[3], a[4], a[5]) for a in O.conv_specs(cn)]


def test_model_requires_gpu():
    import torch
    from yolov3_tensorflow_b200 import yolov3, _lib
    from oracle import yolov3_oracle as O
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.YoloB200Error):
        yolov3(80, O.COCO_ANCHORS)
