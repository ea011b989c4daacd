python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BQ_ARGS > gpurun_out/b.json 2>gpurun_out/b.err; python - <<PY
import json
d=json.load(open("gpurun_out/b.json"))
print(d["value"]/1e9, d["ms_per_step"], d["hbm_roofline_frac_96B"], d["e2e"]["value"]/1e9, d.get("also"))
for k,v in d["kernel_breakdown"].items(): print("   ",k, round(v["ms_per_epdiff_step"],4), v["launches_per_shoot"], round(v.get("frac",0),3))
PY
