"""Phase ablation of the tensor-path stem kernel (YB_STEM_DBG bitmask: 1 no halo load, 2 no im2col, 4 no MMA, 8 no stores)
and pair-vs-1-CTA choice on the 128-wide layers.  CUDA events, L2 flushed between runs."""
import ctypes as C, os, sys, subprocess
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from yolov3_tensorflow_b200 import _lib as L
n, h, w = 64, 416, 416
x = torch.rand((n, h, w, 3), device="cuda")
wt = (torch.randn((32, 27), device="cuda") * 0.1)
sc = torch.ones(32, device="cuda"); sh = torch.zeros(32, device="cuda")
out = torch.empty((n, h, w, 32), device="cuda", dtype=torch.float16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dbg in (0, 1, 2, 4, 8, 3, 12, 7, 15):
    os.environ["YB_STEM_DBG"] = str(dbg)
    ts = []
    for i in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.check(L.lib.yb_stem_conv_fwd_tc(L.ptr(x), L.ptr(wt), L.ptr(sc), L.ptr(sh), n, h, w, 0, 1, L.ptr(out), L.stream_handle()), "stem")
        b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"stem dbg={dbg:2d}: median {ts[len(ts)//2]:.1f} us")
os.environ["YB_STEM_DBG"] = "0"
for shape in ("64 104 104 64 128 3 1", "64 52 52 256 128 1 1", "64 208 208 64 128 3 2"):
    for mode in ("1cta", "2cta"):
        env = dict(os.environ, YB_CONV_MODE=mode)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "conv_probe.py")] + shape.split() + ["6"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
