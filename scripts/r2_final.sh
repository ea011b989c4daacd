#!/bin/bash
# final captures of round 2 on one B200: sanitizers, profiles, bench lines
mkdir -p gpurun_out
for t in memcheck racecheck synccheck; do (timeout 500 compute-sanitizer --tool $t --print-limit 20 python scripts/sanitize_targets.py 2>&1 | tail -50) > gpurun_out/r2_sanitizer_$t.log; echo "== $t: $(tail -1 gpurun_out/r2_sanitizer_$t.log)"; done
bash scripts/r2_profile.sh > /dev/null 2>&1
python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python scripts/mixed_bench.py > gpurun_out/r2_mixed_fft.log 2>&1
python bench_ops.py > gpurun_out/r2_ops.json 2> gpurun_out/r2_ops.err
python - <<'PY'
import json
for f in ("r2_bench_c2", "r2_bench_c3", "r2_bench_reference"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"] / 1e9, d.get("hbm_roofline_frac_96B"), (d.get("e2e") or {}).get("value", 0) / 1e9, (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out | head -30
