# ablation of the 1-CTA conv kernel (YB_CONV_DBG: 1 no A loads, 2 no B loads, 4 no MMA, 8 no epilogue)
for shape in "64 104 104 64 128 3 1" "64 104 104 128 64 1 1" "64 52 52 256 128 1 1" "64 208 208 32 64 3 2"; do
for br in 0 1; do for dbg in 0 1 3 4 8 12 9; do YB_CONV_BRES=$br YB_CONV_DBG=$dbg python tools/conv_probe.py $shape 6 2>&1 | tail -1 | sed "s/^/bres=$br /"; done; done; done
python -m pytest tests/test_gpu_train_ops.py -m gpu -q 2>&1 | tail -3
