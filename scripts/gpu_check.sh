(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --steps 3 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<PY
import json
for f in ("bench_c2","bench_c3"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, d["value"]/1e9, d["ms_per_step"], d["hbm_roofline_frac_96B"], d["e2e"]["value"]/1e9 if d.get("e2e") else None, d.get("also"), d["clocks"], d["roofline"]["frac"], d["roofline"]["traffic"])
    for k,v in d["kernel_breakdown"].items(): print("   ",k, round(v["ms_per_epdiff_step"],4), v["launches_per_shoot"], round(v.get("frac",0),3))
print(open("gpurun_out/bench_reference.json").read()[:300])
PY
