"""In-kernel timeline of one conv launch: CTA 0 stamps clock64 at its pipeline events (yb_debug_set_conv_trace).
Usage: conv_trace.py n h w cin cout k s [res]   (options through YB_CONV_* as for conv_probe.py)
Prints, per tile iteration of CTA 0, cycles relative to kernel entry for: producer (start/end of issue), MMA thread
(tempty wait start/end, first full barrier, last commit) and epilogue warp 2 (tfull wait start/end, per chunk:
before tcgen05.wait::ld / after it / after the TMA store issue, tile end)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from yolov3_tensorflow_b200 import _lib as L
n, h, w, cin, cout, k, s = (int(v) for v in sys.argv[1:8])
with_res = len(sys.argv) > 8 and sys.argv[8] == "res"
dev = "cuda"
x = torch.randn((n, h, w, cin), device=dev).half()
cp = L.lib.yb_conv_cout_pad(cout)
wp = (torch.randn((cp, k, k, cin), device=dev) * 0.05).half()
sc = torch.ones(cp, device=dev); sh = torch.zeros(cp, device=dev)
out = torch.empty((n, h // s, w // s, cout), device=dev, dtype=torch.float16)
res = torch.randn((n, h // s, w // s, cout), device=dev).half() if with_res else None
d = L.ConvDesc(n=n, h=h, w=w, cin=cin, cout=cout, ksize=k, stride=s, in_ld=cin, out_ld=cout, res_ld=cout, dtype=0, out_fp32=0, leaky=1, upsample2x=0)
tr = torch.zeros(10 * 64 * 32 + 2, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    L.check(L.lib.yb_conv2d_fwd(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(sc), L.ptr(sh), L.ptr(res), L.ptr(out), None, None, L.stream_handle()), "conv")
for _ in range(3):
    run()
flush.zero_()
L.check(L.lib.yb_debug_set_conv_trace(L.ptr(tr)), "trace")
run(); torch.cuda.synchronize()
L.check(L.lib.yb_debug_set_conv_trace(None), "trace")
t = tr.cpu().numpy()
t0 = int(t[-2]); body = t[:-2].reshape(10, 64, 32)
opts = " ".join(f"{k_[3:]}={L.get_option(k_)}" for k_ in ("YB_CONV_DBG", "YB_CONV_MODE", "YB_CONV_EPI", "YB_CONV_KPS", "YB_CONV_EG") if L.get_option(k_))
print(f"# [{opts}] n{n} {h}x{w} {cin}->{cout} k{k}s{s}{' +res' if with_res else ''}; cycles since kernel entry; set-up done at {int(t[-1]) - t0}")
def rel(v): return "-" if v == 0 else str(int(v) - t0)
print("it | prod start end | mma: wait_tempty got_tempty first_full last_commit | epi(w2): wait_tfull got_tfull [ch: pre_ldwait post_ldwait store]... end")
for it in range(64):
    if body[0, it, 0] == 0 and body[1, it, 0] == 0 and body[2, it, 0] == 0:
        break
    pr = " ".join(rel(body[0, it, j]) for j in range(2))
    mm = " ".join(rel(body[1, it, j]) for j in range(4))
    ep = " ".join(rel(body[2, it, j]) for j in range(2))
    chs = []
    for ch in range(8):
        if body[2, it, 2 + 3 * ch] == 0: break
        chs.append("[" + " ".join(rel(body[2, it, 2 + 3 * ch + j]) for j in range(3)) + "]")
    fine = " fine(ch0: ffma leaky sts fence)=" + "/".join(rel(body[2, it, j]) for j in (20, 21, 22, 23)) if body[2, it, 20] else ""
    print(f"{it:2d} | {pr} | {mm} | {ep} {' '.join(chs)} {rel(body[2, it, 31])}{fine}")
for it in range(64):          # second epilogue group (warp 6), if it ran
    if body[6, it, 0] == 0: continue
    print(f"{it:2d} | epi(w6): {rel(body[6, it, 0])} {rel(body[6, it, 1])} ... {rel(body[6, it, 31])}")
last = max(int(body[r, :, :].max()) for r in range(10))
print(f"# last stamp at {last - t0} cycles")
