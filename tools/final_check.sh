timeout 110 python -m pytest tests/test_gpu_path.py tests/test_gpu_train_ops.py tests/test_gpu_conv.py -m gpu -x -q 2>&1 | tail -2
timeout 80 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_kps.json 2> gpurun_out/bench_kps.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_kps.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["train"]["ms_per_step"])
PY
