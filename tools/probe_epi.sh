for shape in "64 104 104 64 128 3 1" "64 104 104 128 64 1 1" "64 52 52 256 128 1 1" "64 26 26 256 512 3 1" "64 52 52 128 256 3 1"; do python tools/conv_probe.py $shape 6 2>&1 | tail -1; done
