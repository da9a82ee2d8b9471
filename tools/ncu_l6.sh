# one full ncu capture of the 1-CTA conv kernel on the 104x104 64->128 3x3 layer
ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel --launch-skip 2 --launch-count 1 \
    -o gpurun_out/l6_1cta -f python tools/conv_probe.py 64 104 104 64 128 3 1 2 > gpurun_out/ncu_l6.log 2>&1
for dbg in 7 11 3; do YB_CONV_DBG=$dbg python tools/conv_probe.py 64 104 104 64 128 3 1 6 2>&1 | tail -1; done
