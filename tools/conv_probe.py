"""Time one conv shape through yb_conv2d_fwd (CUDA events, L2 flushed between runs) under the current
YB_CONV_* options.  Usage: conv_probe.py n h w cin cout k s [iters] [res]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from yolov3_tensorflow_b200 import _lib as L
n, h, w, cin, cout, k, s = (int(v) for v in sys.argv[1:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 10
with_res = len(sys.argv) > 9 and sys.argv[9] == "res"
dev = "cuda"
x = torch.randn((n, h, w, cin), device=dev).half()
cp = L.lib.yb_conv_cout_pad(cout)
wp = (torch.randn((cp, k, k, cin), device=dev) * 0.05).half()
sc = torch.ones(cp, device=dev); sh = torch.zeros(cp, device=dev)
out = torch.empty((n, h // s, w // s, cout), device=dev, dtype=torch.float16)
res = torch.randn((n, h // s, w // s, cout), device=dev).half() if with_res else None
d = L.ConvDesc(n=n, h=h, w=w, cin=cin, cout=cout, ksize=k, stride=s, in_ld=cin, out_ld=cout, res_ld=cout, dtype=0, out_fp32=0, leaky=1, upsample2x=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(iters + 2):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if os.environ.get("YB_PROBE_HALO") == "1":
        L.check(L.lib.yb_conv3x3_halo_fwd(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(sc), L.ptr(sh), L.ptr(res), L.ptr(out), L.stream_handle()), "conv_halo")
    else:
        L.check(L.lib.yb_conv2d_fwd(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(sc), L.ptr(sh), L.ptr(res), L.ptr(out), None, None, L.stream_handle()), "conv")
    b.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
fl = 2.0 * n * (h // s) * (w // s) * cout * cin * k * k
byt = 2.0 * n * (h * w * cin + (h // s) * (w // s) * cout * (2 if with_res else 1))
opts = ("HALO " if os.environ.get("YB_PROBE_HALO") == "1" else "") + " ".join(f"{k_[3:]}={L.get_option(k_)}" for k_ in ("YB_CONV_DBG", "YB_CONV_MODE", "YB_CONV_EPI", "YB_CONV_BRES", "YB_CONV_KPS", "YB_CONV_EG") if L.get_option(k_))
print(f"[{opts}] n{n} {h}x{w} {cin}->{cout} k{k}s{s}{' +res' if with_res else ''}: median {ts[len(ts)//2]:.1f} us  min {ts[0]:.1f} us  "
      f"{fl/ts[len(ts)//2]/1e6:.0f} TFLOP/s  {byt/ts[len(ts)//2]/1e3:.0f} GB/s(alg)")
