"""CPU test of bench.py's reference arm (the CPU oracle port): one JSON line with the contract's keys.  The GPU arm is
exercised on the GPU box; this guards the part of the contract that must also work without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, YB_CPU_THREADS=str(min(8, os.cpu_count() or 1)), YB_CPU_IMAGES="2", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["product_modules_loaded"] == []          # the reference arm never touches the product package / its .so
    assert d["steps"] >= 3 and len(cb["candidates"]) >= 1 and "images_per_s" in cb["candidates"][-1]


def test_bench_params_follow_the_products_conv_table():
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from yolov3_tensorflow_b200.model import yolov3
    ps = bench.make_bench_params()
    table = yolov3.conv_table(80)
    assert len(ps) == len(table) == 75
    for p, (cin, cout, k, s, bn) in zip(ps, table):
        assert p["w"].shape == (k, k, cin, cout) and (("gamma" in p) == bool(bn))
    assert np.all(ps[58]["b"].reshape(3, -1)[:, 4] == -2.0)
    # the GPU arm builds the same parameters from the product's table (it never imports oracle/): identical arrays
    ps2 = bench.make_bench_params(specs=table)
    for p, q in zip(ps, ps2):
        assert p.keys() == q.keys() and all(np.array_equal(p[k], q[k]) for k in p)
