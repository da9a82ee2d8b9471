"""Run a few training steps (bench.py's training configuration) — target for ncu launch lists."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench, yolov3_tensorflow_b200 as pkg
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ts = int(sys.argv[2]) if len(sys.argv) > 2 else 416
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
anchors = pkg.parse_anchors(os.path.join(ROOT, "yolov3_tensorflow_b200", "data", "yolo_anchors.txt"))
m = pkg.yolov3(80, anchors, use_label_smooth=True, use_focal_loss=True, batch_norm_decay=0.99, dtype="bf16")
m.init_params(3)
rng = np.random.default_rng(3)
x = torch.from_numpy(rng.random((tb, ts, ts, 3), dtype=np.float32)).cuda()
y = bench.synth_y_true(rng, tb, ts, anchors)
m.train_step(x, y, 1e-4)   # warm-up: builds the plan, packs weights
torch.cuda.synchronize(); torch.cuda.profiler.start()   # ncu --profile-from-start off: only the steps are captured
for _ in range(steps):
    l = m.train_step(x, y, 1e-4)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("loss", float(l[0]))
