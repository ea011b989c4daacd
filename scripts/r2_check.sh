#!/bin/bash
# round-2 check on one B200: GPU test-suite (all failures listed), smoke, bench lines
mkdir -p gpurun_out
(time timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2_pytest.log 2>&1
tail -25 gpurun_out/r2_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
(time python bench.py) > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
tail -5 gpurun_out/r2_bench_c2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c2.json"))
    print("C2", d["value"] / 1e9, d["ms_per_step"], d["hbm_roofline_frac_96B"], "e2e", d["e2e"]["value"] / 1e9, d["clocks"])
    for k, v in d["kernel_breakdown"].items():
        print("   ", k, round(v["ms_per_epdiff_step"], 4), round(v.get("frac", 0), 3))
    print("cpu", d["cpu_baseline"])
    for k, v in d.get("also", {}).items():
        print("also", k, json.dumps(v)[:600])
except Exception as e:
    print("ERR", e, open("gpurun_out/r2_bench_c2.json").read()[:500])
PY
