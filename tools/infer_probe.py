"""Run a few inference steps (forward + decode + batched NMS, bench.py's configuration) — target for ncu."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import yolov3_tensorflow_b200 as pkg
from yolov3_tensorflow_b200.utils.nms_utils import batched_gpu_nms
b = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
anchors = pkg.parse_anchors(os.path.join(ROOT, "yolov3_tensorflow_b200", "data", "yolo_anchors.txt"))
m = pkg.yolov3(80, anchors, dtype="fp16")
m.init_params(3)
rng = np.random.default_rng(3)
x = torch.from_numpy(rng.random((b, size, size, 3), dtype=np.float32)).cuda()
fms = m.forward(x, is_training=False); b_, s_ = m.predict_scores(fms); batched_gpu_nms(b_, s_, 80, max_boxes=200, score_thresh=0.3, nms_thresh=0.45)   # warm-up
torch.cuda.synchronize(); torch.cuda.profiler.start()   # ncu --profile-from-start off: only the steps are captured
for _ in range(steps):
    fms = m.forward(x, is_training=False)
    boxes, scores = m.predict_scores(fms)
    out = batched_gpu_nms(boxes, scores, 80, max_boxes=200, score_thresh=0.3, nms_thresh=0.45)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("detections", sum(int(o[0].shape[0]) for o in out))
