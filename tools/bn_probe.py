"""Time the BN streaming kernels (csrc/bn.cu) on the training step's layer shapes (batch 32 @416, bf16): CUDA events,
L2 flushed between runs, median of `iters`.  Prints us and GB/s of algorithmic traffic per kernel and kernel shape
(YB_BN_CPT = 4 | 8).  Usage: bn_probe.py [iters]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from yolov3_tensorflow_b200 import _lib as L
lib, ptr, st = L.lib, L.ptr, L.stream_handle
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 7
n = 32
shapes = [(416, 32), (208, 64), (208, 32), (104, 128), (104, 64), (52, 256), (52, 128), (26, 512), (26, 256), (13, 1024), (13, 512)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
need = C.c_size_t(); L.check(lib.yb_bn_bwd_reduce_workspace_bytes(C.byref(need)), "ws")
ws = torch.zeros(need.value, dtype=torch.uint8, device="cuda")

def timed(fn):
    ts = []
    for i in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

print("| HxW | c | MB/tensor | kernel | cpt4 us | GB/s | cpt8 us | GB/s |")
print("|--:|--:|--:|---|--:|--:|--:|--:|")
tot = {"4": 0.0, "8": 0.0}
for hw, c in shapes:
    rows = n * hw * hw
    z = torch.randn((rows, c), device="cuda").bfloat16()
    dA = (torch.randn((rows, c), device="cuda") * 0.1).bfloat16()
    out = torch.empty_like(z)
    gamma = torch.rand(c, device="cuda") + 0.5; beta = torch.zeros(c, device="cuda")
    ssum = z.float().sum(0); ssq = (z.float() ** 2).sum(0)
    mm = torch.zeros(c, device="cuda"); mv = torch.ones(c, device="cuda")
    sc = torch.empty(c, device="cuda"); sh = torch.empty(c, device="cuda"); sm = torch.empty(c, device="cuda"); si = torch.empty(c, device="cuda")
    dg = torch.empty(c, device="cuda"); db = torch.empty(c, device="cuda")
    mb = rows * c * 2 / 1e6
    kernels = {
        "stats_act_apply": (2, lambda: L.check(lib.yb_bn_stats_act_apply(ptr(z), c, ptr(ssum), ptr(ssq), ptr(gamma), ptr(beta), 1e-5, 0.99, ptr(mm), ptr(mv),
                                                ptr(sc), ptr(sh), ptr(sm), ptr(si), None, 0, ptr(out), c, n, hw, hw, c, L.YB_BF16, 1, 0, st()), "act")),
        "bwd_reduce": (2, lambda: L.check(lib.yb_bn_bwd_reduce(ptr(dA), c, ptr(z), c, ptr(sc), ptr(sh), ptr(sm), ptr(si), n, hw, hw, c, L.YB_BF16, 1, 0,
                                           ptr(dg), ptr(db), ptr(ws), st()), "red")),
        "bwd_apply": (3, lambda: L.check(lib.yb_bn_bwd_apply(ptr(dA), c, ptr(z), c, ptr(gamma), ptr(sc), ptr(sh), ptr(sm), ptr(si), ptr(dg), ptr(db),
                                          n, hw, hw, c, L.YB_BF16, 1, 0, 0, ptr(out), c, st()), "app")),
    }
    for name, (units, fn) in kernels.items():
        res = []
        for cpt in ("4", "8"):
            L.set_option("YB_BN_CPT", None if cpt == "8" else "4")
            us = timed(fn); tot[cpt] += us
            res.append(f"{us:.1f} | {units * mb / us * 1e3:.0f}")
        L.set_option("YB_BN_CPT", None)
        print(f"| {hw} | {c} | {mb:.1f} | {name} | {res[0]} | {res[1]} |", flush=True)
print("sum over the listed shapes: cpt4 %.1f us, cpt8 %.1f us" % (tot["4"], tot["8"]))
