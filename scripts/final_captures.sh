#!/bin/bash
# final captures of round 2 on one B200: full GPU suite, sanitizers, profiles, bench lines (profiles/r2_*)
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_pytest.log 2>&1; tail -4 gpurun_out/r2_pytest.log
for t in memcheck racecheck synccheck; do (timeout 600 compute-sanitizer --tool $t --print-limit 20 python scripts/sanitize_targets.py 2>&1 | tail -60) > gpurun_out/r2_sanitizer_$t.log; echo "== $t: $(tail -1 gpurun_out/r2_sanitizer_$t.log)"; done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1
K='regex:gather3|adstar_ring|xpass|slab|compose_ring'
ncu --set full --clock-control none --import-source on -k "$K" -s 8 -c 5 -o /tmp/r2_c2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 8 -c 5 -o /tmp/r2_c3_full -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_c3.log 2>&1
python scripts/profile_summary.py /tmp/r2_c2_full.ncu-rep gpurun_out/r2_ncu_full_summary.txt gpurun_out/r2_traffic.json "C2: 16 x 128^3" "ncu --set full --clock-control none --import-source on -k '$K' -s 8 -c 5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
python scripts/profile_summary.py /tmp/r2_c3_full.ncu-rep gpurun_out/r2_ncu_full_summary_c3.txt gpurun_out/r2_traffic_c3.json "C3 share: 8 x 256^3" "ncu --set full ... python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
rm -f gpurun_out/ncu_*.log
( time python bench.py ) > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
( time python bench.py --impl reference ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python bench_ops.py > gpurun_out/r2_ops.json 2> gpurun_out/r2_ops.err
python scripts/atlas_profile.py > gpurun_out/r2_atlas_profile.txt 2>&1
python - <<'PY'
import json
for f in ("r2_bench_c2", "r2_bench_c3", "r2_bench_reference"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"] / 1e9, d.get("hbm_roofline_frac_96B"), (d.get("e2e") or {}).get("value", 0) / 1e9, (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -4 gpurun_out/r2_bench_c2.err
