#!/bin/bash
# round-2 GPU session 14: fused stem + Conv_1
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider --tb=short -x -k "stem_conv1_fused" 2>&1 | tail -15 | cut -c1-300
python - <<'PY'
import ctypes as C, torch, sys
sys.path.insert(0, ".")
from yolov3_tensorflow_b200 import _lib as L
n, h, w = 64, 416, 416
x = torch.rand((n, h, w, 3), device="cuda")
w0 = torch.randn((32, 3, 3, 3), device="cuda") / 5; s0 = torch.ones(32, device="cuda"); b0 = torch.zeros(32, device="cuda")
w1p = (torch.randn((64, 3, 3, 32), device="cuda") / 17).half(); s1 = torch.ones(64, device="cuda"); b1 = torch.zeros(64, device="cuda")
d = L.ConvDesc(n=n, h=h, w=w, cin=32, cout=64, ksize=3, stride=2, in_ld=32, out_ld=64, res_ld=0, dtype=0, out_fp32=0, leaky=1, upsample2x=0)
a0 = torch.empty((n, h, w, 32), dtype=torch.float16, device="cuda")
out = torch.empty((n, h // 2, w // 2, 64), dtype=torch.float16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, name):
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); print(f"{name}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us")
st = L.stream_handle
timeit(lambda: L.check(L.lib.yb_stem_conv_fwd_tc(L.ptr(x), L.ptr(w0), L.ptr(s0), L.ptr(b0), n, h, w, 0, 1, L.ptr(a0), st()), "stem"), "stem alone")
timeit(lambda: L.check(L.lib.yb_conv3x3_halo_fwd(C.byref(d), L.ptr(a0), L.ptr(w1p), L.ptr(s1), L.ptr(b1), None, L.ptr(out), st()), "halo"), "Conv_1 halo alone")
timeit(lambda: L.check(L.lib.yb_stem_conv1_fused_fwd(C.byref(d), L.ptr(x), L.ptr(w0), L.ptr(s0), L.ptr(b0), L.ptr(w1p), L.ptr(s1), L.ptr(b1), L.ptr(out), st()), "fused"), "stem + Conv_1 fused")
PY
