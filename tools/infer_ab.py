"""Time the inference step (bench.py's configuration: detect_raw, batch 64 @416, fp16, cfg-2 weights) under sets of
runtime options; a fresh model (and plan) per set, because most conv options are read when the plan is bound.
CUDA events, L2 flushed between steps, median of `iters`.  Usage: infer_ab.py [batch] [size] [iters] -- "K=V,K=V" "" ..."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import yolov3_tensorflow_b200 as pkg
from yolov3_tensorflow_b200 import _lib
from bench import make_bench_params, NMS_ARGS
args = sys.argv[1:]
sets = [""]
if "--" in args:
    k = args.index("--"); sets = args[k + 1:]; args = args[:k]
b = int(args[0]) if len(args) > 0 else 64
size = int(args[1]) if len(args) > 1 else 416
iters = int(args[2]) if len(args) > 2 else 10
anchors = pkg.parse_anchors(os.path.join(ROOT, "yolov3_tensorflow_b200", "data", "yolo_anchors.txt"))
params = make_bench_params()
x = torch.from_numpy(np.random.default_rng(2).random((b, size, size, 3), dtype=np.float32)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for spec in sets:
    kv = [s.split("=") for s in spec.split(",") if s]
    for k_, v_ in kv: _lib.set_option(k_, v_)
    try:
        m = pkg.yolov3(80, anchors, dtype="fp16")
        m.set_params(params, "HWIO")
        for _ in range(3): out = m.detect_raw(x, **NMS_ARGS)
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = m.detect_raw(x, **NMS_ARGS); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        ts.sort()
        med = ts[len(ts) // 2]
        print(f"[{spec or 'default'}] batch {b} @{size}: median {med:.3f} ms ({b / med * 1e3:.0f} img/s)  min {ts[0]:.3f} ms  detections {int(out[5].sum())}", flush=True)
        del m
    finally:
        for k_, _ in kv: _lib.set_option(k_, None)
