# round-1 final captures: launch list, full-set reports (c2, c3), bench lines
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gather3|xpass2|slab" -s 5 -c 5 -o gpurun_out/c2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gather3|xpass2|slab" -s 5 -c 5 -o gpurun_out/c3_full -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_c3.log 2>&1
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python -c "
import json
for f in ('bench_c2','bench_c3','bench_reference'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value']/1e9, d.get('hbm_roofline_frac_96B'), (d.get('e2e') or {}).get('value',0)/1e9, d.get('roofline'))
"
