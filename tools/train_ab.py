"""Time the training step (bench.py's configuration: batch 32 @416, bf16) under sets of runtime options, CUDA events,
median of `iters` steps after warm-up.  Usage: train_ab.py [batch] [size] [iters] -- "K=V,K=V" "K=V" ...
An empty set "" is the default configuration.  YB_TRAIN_PHASE=fwd times the forward + loss only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, yolov3_tensorflow_b200 as pkg
from yolov3_tensorflow_b200 import _lib
args = sys.argv[1:]
sets = [""]
if "--" in args:
    k = args.index("--"); sets = args[k + 1:]; args = args[:k]
tb = int(args[0]) if len(args) > 0 else 32
ts = int(args[1]) if len(args) > 1 else 416
iters = int(args[2]) if len(args) > 2 else 10
anchors = pkg.parse_anchors(os.path.join(ROOT, "yolov3_tensorflow_b200", "data", "yolo_anchors.txt"))
rng = np.random.default_rng(3)
x = torch.from_numpy(rng.random((tb, ts, ts, 3), dtype=np.float32)).cuda()
y = bench.synth_y_true(rng, tb, ts, anchors)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for spec in sets:
    kv = [s.split("=") for s in spec.split(",") if s]
    for k_, v_ in kv: _lib.set_option(k_, v_)
    try:
        # a fresh model (and plan) per set: most conv options are read when the plan is bound
        m = pkg.yolov3(80, anchors, use_label_smooth=True, use_focal_loss=True, batch_norm_decay=0.99, dtype="bf16")
        m.init_params(3)
        for _ in range(3): m.train_step(x, y, 1e-4)
        ts_ = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); l = m.train_step(x, y, 1e-4); b.record(); torch.cuda.synchronize()
            ts_.append(a.elapsed_time(b))
        ts_.sort()
        print(f"[{spec or 'default'}] batch {tb} @{ts}: median {ts_[len(ts_)//2]:.3f} ms  min {ts_[0]:.3f} ms  loss {float(l[0]):.4f}", flush=True)
        del m
    finally:
        for k_, _ in kv: _lib.set_option(k_, None)
