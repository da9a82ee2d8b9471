# per-kernel evidence: duration, DRAM bytes, tensor-pipe and SM utilisation for one inference step and one training step
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/kern_infer.csv python tools/infer_probe.py 64 416 1 > gpurun_out/kern_infer.log 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/kern_train.csv python tools/train_probe.py 32 416 1 > gpurun_out/kern_train.log 2>&1
tail -n 2 gpurun_out/kern_infer.log; tail -n 2 gpurun_out/kern_train.log
