#!/bin/bash
# memcheck / racecheck over the kernels added for the training path (BN streaming kernels, one-launch repack, stem with
# statistics, side-stream wgrad) + one ncu --set full capture of the wgrad kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SEL='bn_train_forward_backward or col_sum or dgrad_weight_repack or train_step_frozen_bn or stem_conv_tensor_core or wgrad_matches'
echo "=== memcheck"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_path.py tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|=========     at " | head -12 > gpurun_out/r02_s_memcheck.txt; cat gpurun_out/r02_s_memcheck.txt | cut -c1-200
echo "=== racecheck (BN kernels, repack, stem)"
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_conv.py -m gpu -q -p no:cacheprovider -x -k "bn_train_forward_backward or stem_conv_tensor_core" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|=========     at " | head -12 > gpurun_out/r02_s_racecheck.txt; cat gpurun_out/r02_s_racecheck.txt | cut -c1-200
echo "=== ncu --set full: wgrad 3x3 128->256 @52x52, batch 32"
cat > /tmp/wg_probe.py <<'PY'
import ctypes as C, os, sys, torch
sys.path.insert(0, os.getcwd())
from yolov3_tensorflow_b200 import _lib as L
n, h, w, cin, cout, k = 32, 52, 52, 128, 256, 3
x = torch.randn((n, h, w, cin), device="cuda").bfloat16(); dz = (torch.randn((n, h, w, cout), device="cuda") * 0.1).bfloat16()
dw = torch.zeros((cout, k, k, cin), device="cuda")
d = L.ConvDesc(n=n, h=h, w=w, cin=cin, cout=cout, ksize=k, stride=1, in_ld=cin, out_ld=cout, res_ld=0, dtype=1, out_fp32=0, leaky=0, upsample2x=0)
for _ in range(2):
    L.check(L.lib.yb_conv2d_wgrad(C.byref(d), L.ptr(x), L.ptr(dz), cout, 0, L.ptr(dw), L.stream_handle()), "wgrad")
torch.cuda.synchronize(); print("ok", float(dw.abs().mean()))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -c 1 -f -o gpurun_out/r02_s_ncu_full_wgrad python /tmp/wg_probe.py > gpurun_out/r02_s_ncu_full.log 2>&1; tail -2 gpurun_out/r02_s_ncu_full.log | cut -c1-200
