"""Time one conv shape through yb_conv2d_fwd (CUDA events, L2 flushed between runs) under the current
YB_CONV_* environment.  Usage: conv_probe.py n h w cin cout k s [iters]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from yolov3_tensorflow_b200 import _lib as L
n, h, w, cin, cout, k, s = (int(v) for v in sys.argv[1:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 10
dev = "cuda"
x = torch.randn((n, h, w, cin), device=dev).half()
cp = L.lib.yb_conv_cout_pad(cout)
wp = (torch.randn((cp, k, k, cin), device=dev) * 0.05).half()
sc = torch.ones(cp, device=dev); sh = torch.zeros(cp, device=dev)
out = torch.empty((n, h // s, w // s, cout), device=dev, dtype=torch.float16)
d = L.ConvDesc(n=n, h=h, w=w, cin=cin, cout=cout, ksize=k, stride=s, in_ld=cin, out_ld=cout, res_ld=0, dtype=0, out_fp32=0, leaky=1, upsample2x=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(iters + 2):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(L.lib.yb_conv2d_fwd(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(sc), L.ptr(sh), None, L.ptr(out), None, None, L.stream_handle()), "conv")
    b.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
fl = 2.0 * n * (h // s) * (w // s) * cout * cin * k * k
print(f"dbg={os.environ.get('YB_CONV_DBG','0')} mode={os.environ.get('YB_CONV_MODE','auto')} mc={os.environ.get('YB_CONV_MC','-')} "
      f"shape n{n} {h}x{w} {cin}->{cout} k{k}s{s}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  {fl/ts[len(ts)//2]/1e6:.0f} TFLOP/s")
