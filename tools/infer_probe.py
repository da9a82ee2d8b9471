"""Run a few inference steps exactly as bench.py does (model.detect_raw: forward + decode + NMS in one engine call,
bench.py's cfg-2 weights) between cudaProfilerStart/Stop — the target of the ncu per-kernel captures."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import yolov3_tensorflow_b200 as pkg
from bench import make_bench_params, NMS_ARGS
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
anchors = pkg.parse_anchors(os.path.join(ROOT, "yolov3_tensorflow_b200", "data", "yolo_anchors.txt"))
m = pkg.yolov3(80, anchors, dtype="fp16")
m.set_params(make_bench_params(), "HWIO")
x = torch.from_numpy(np.random.default_rng(2).random((b, size, size, 3), dtype=np.float32)).cuda()
for _ in range(2):
    out = m.detect_raw(x, **NMS_ARGS)          # warm-up
torch.cuda.synchronize(); torch.cuda.profiler.start()   # ncu --profile-from-start off: only the steps are captured
for _ in range(steps):
    out = m.detect_raw(x, **NMS_ARGS)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("detections", int(out[5].sum()))
