#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -q -p no:cacheprovider --tb=short -x -s -k "test_train_step_matches_oracle and flags1" > gpurun_out/r02_p_run$i.txt 2>&1
  tail -1 gpurun_out/r02_p_run$i.txt
done
python - <<'PY'
import re,glob
worst={}
for f in sorted(glob.glob("gpurun_out/r02_p_run*.txt")):
    for l in open(f):
        m=re.match(r"\s+(\d+) (\w+)\s+\| ([\d.e+-]+) \| ([\d.e+-]+)", l)
        if m:
            i,k,e,n=int(m.group(1)),m.group(2),float(m.group(3)),float(m.group(4))
            r=e/(3*n+0.02)
            key=(i,k)
            if key not in worst or r>worst[key][0]: worst[key]=(r,e,n,f[-8:])
top=sorted(worst.items(), key=lambda kv:-kv[1][0])[:10]
for k,v in top: print(k, "ratio to bar %.3f  e %.4g noise %.4g  %s"%v)
PY
grep -h "^E  " gpurun_out/r02_p_run*.txt | cut -c1-600 | head -5
