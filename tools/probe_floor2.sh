for dbg in 15 12 8; do YB_CONV_DBG=$dbg python tools/conv_probe.py 64 104 104 64 128 3 1 6 2>&1 | tail -1; done
for dbg in 15 7; do YB_CONV_DBG=$dbg python tools/conv_probe.py 64 52 52 256 128 1 1 6 2>&1 | tail -1; done
