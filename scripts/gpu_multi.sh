N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c2_${N}gpu.json 2> gpurun_out/bench_c2_${N}gpu.err
python bench_atlas.py --size 256 --subjects-per-gpu 8 --batch 8 --steps 2 --warmup 1 > gpurun_out/atlas_c3_1gpu.json 2> gpurun_out/atlas_1gpu.err
$TR bench_atlas.py --size 256 --subjects-per-gpu 8 --batch 8 --steps 2 --warmup 1 > gpurun_out/atlas_c3_${N}gpu.json 2> gpurun_out/atlas_${N}gpu.err
tail -1 gpurun_out/bench_c2_${N}gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['n_gpus'], d['value']/1e9, d['hbm_roofline_frac_96B'], d['e2e']['value']/1e9)"
tail -1 gpurun_out/atlas_c3_1gpu.json; tail -1 gpurun_out/atlas_c3_${N}gpu.json
