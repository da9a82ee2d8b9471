python tools/conv_probe.py 64 208 208 32 64 3 1 6 2>&1 | tail -1
python tools/conv_probe.py 64 416 416 32 64 3 2 6 2>&1 | tail -1
YB_THIN=0 python bench.py --no-cpu-baseline --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('YB_THIN=0', d['value'], d['ms_per_step'], d['roofline']['frac'])"
