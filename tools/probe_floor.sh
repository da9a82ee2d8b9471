for shape in "64 26 26 256 512 3 1" "64 104 104 64 128 3 1"; do for dbg in 0 3 4 7 11; do YB_CONV_DBG=$dbg python tools/conv_probe.py $shape 6 2>&1 | tail -1; done; done
