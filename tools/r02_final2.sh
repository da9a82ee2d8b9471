#!/bin/bash
# round-2 final evidence (second pass, after the training-path work): all GPU tests, bench (with the CPU baseline), ncu launch list of
# the bench command, per-kernel ncu tables of one inference step, one ncu --set full capture of the top kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== tests"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | tail -15 > gpurun_out/r02_r_tests.log; tail -4 gpurun_out/r02_r_tests.log | cut -c1-300
echo "=== bench"
timeout 1500 python bench.py > gpurun_out/r02_r_bench.json 2> gpurun_out/r02_r_bench.err; tail -c 400 gpurun_out/r02_r_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_r_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]; c=d.get("cpu_baseline",{})
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof", r["frac"], "conv ms", r["ms_per_step_conv"], "nms", r["ms_per_step_nms"], "train", d["train"]["ms_per_step"], d["train"]["e2e"]["ms_per_step"], "train608", d.get("train608",{}).get("ms_per_step"), "lat", d["latency_batch1"]["ms_median"], d["latency_batch1"]["cuda_graph_ms_median"], "cpu", c.get("value"), "clocks", d["clocks"])
except Exception as e: print("bench parse failed", e)
PY
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_r_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/r02_r_launches.log 2>&1; tail -2 gpurun_out/r02_r_launches.log | cut -c1-200
python tools/launch_table.py gpurun_out/r02_r_launches.csv 4 2>&1 | tail -12
echo "=== ncu per-kernel tables (one inference step, one training step)"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_r_kern_infer.csv python tools/infer_probe.py 64 416 1 > gpurun_out/r02_r_kern_infer.log 2>&1; tail -1 gpurun_out/r02_r_kern_infer.log
python tools/layer_table.py gpurun_out/r02_r_kern_infer.csv 64 416 > gpurun_out/r02_r_layers_infer.md 2> gpurun_out/r02_r_conv_traffic.json; tail -3 gpurun_out/r02_r_layers_infer.md; cat gpurun_out/r02_r_conv_traffic.json
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_r_kern_train.csv python tools/train_probe.py 32 416 1 > gpurun_out/r02_r_kern_train.log 2>&1; tail -1 gpurun_out/r02_r_kern_train.log
python tools/kernel_table.py gpurun_out/r02_r_kern_train.csv > gpurun_out/r02_r_kernels_train.md 2>&1; head -30 gpurun_out/r02_r_kernels_train.md | cut -c1-200
python tools/kernel_table.py gpurun_out/r02_r_kern_infer.csv > gpurun_out/r02_r_kernels_infer.md 2>&1
echo "=== ncu --set full: the fused stem + Conv_1 halo kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 1 -f -o gpurun_out/r02_r_ncu_full_stem_halo python tools/stem_fused_probe.py 64 416 416 > gpurun_out/r02_r_ncu_full.log 2>&1; tail -2 gpurun_out/r02_r_ncu_full.log | cut -c1-200
ls -la gpurun_out | grep r02_r
