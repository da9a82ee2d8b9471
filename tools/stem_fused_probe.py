"""Smallest reproduction of the fused stem + Conv_1 kernel (one launch, 2 x 64 x 32 image)."""
import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov3_tensorflow_b200 import _lib as L
n, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2, 64, 32)
x = torch.rand((n, h, w, 3), device="cuda")
w0 = torch.randn((32, 3, 3, 3), device="cuda") / 5; s0 = torch.ones(32, device="cuda"); b0 = torch.zeros(32, device="cuda")
w1p = (torch.randn((64, 3, 3, 32), device="cuda") / 17).half(); s1 = torch.ones(64, device="cuda"); b1 = torch.zeros(64, device="cuda")
d = L.ConvDesc(n=n, h=h, w=w, cin=32, cout=64, ksize=3, stride=2, in_ld=32, out_ld=64, res_ld=0, dtype=0, out_fp32=0, leaky=1, upsample2x=0)
out = torch.empty((n, h // 2, w // 2, 64), dtype=torch.float16, device="cuda")
L.check(L.lib.yb_stem_conv1_fused_fwd(C.byref(d), L.ptr(x), L.ptr(w0), L.ptr(s0), L.ptr(b0), L.ptr(w1p), L.ptr(s1), L.ptr(b1), L.ptr(out), L.stream_handle()), "fused")
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
