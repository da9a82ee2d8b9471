#!/bin/bash
# round-2 GPU session 12 (2 GPUs): DP equivalence test, bucketed vs blocking all-reduce, graphed latency, training e2e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider --tb=short -s 2>&1 | tail -12 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -q -p no:cacheprovider --tb=short -k "graphed" 2>&1 | tail -5 | cut -c1-300
echo "=== bench N=1"
timeout 900 python bench.py --no-cpu-baseline --steps 8 --no-train608 > gpurun_out/r02_12_bench_n1.json 2> gpurun_out/r02_12_bench_n1.err; tail -c 500 gpurun_out/r02_12_bench_n1.err
for mb in 32 0; do
echo "=== bench N=2 bucket $mb"
YB_BUCKET_MB=$mb timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --no-cpu-baseline > gpurun_out/r02_12_bench_n2_mb$mb.json 2> gpurun_out/r02_12_bench_n2.err; tail -c 500 gpurun_out/r02_12_bench_n2.err
done
python - <<PY
import json
for f in ("n1", "n2_mb32", "n2_mb0"):
    try:
        d=json.loads(open(f"gpurun_out/r02_12_bench_{f}.json").read().strip().splitlines()[-1])
        t=d["train"]
        print(f, "infer", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| train ms", round(t["ms_per_step"],3), "img/s", round(t["images_per_s"]), "e2e ms", round(t["e2e"]["ms_per_step"],3), "| lat", d.get("latency_batch1",{}).get("ms_median"), "graph", d.get("latency_batch1",{}).get("cuda_graph_ms_median"))
    except Exception as e: print(f, "parse failed", e)
PY
