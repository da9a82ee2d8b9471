export YB_CONV_KPS=1
timeout 100 python -m pytest tests/test_gpu_conv.py -m gpu -x -q 2>&1 | tail -2
for shape in "64 208 208 32 64 3 1" "64 104 104 64 128 3 1" "64 104 104 128 64 1 1"; do python tools/conv_probe.py $shape 5 2>&1 | tail -1; done
