N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 --no-extra > gpurun_out/bench_c2_${N}gpu.json 2> gpurun_out/bench_c2_${N}gpu.err
tail -1 gpurun_out/bench_c2_${N}gpu.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['n_gpus'], d['value']/1e9, d['hbm_roofline_frac_96B'], d['e2e']['value']/1e9, d['config'].get('host_numa_node'))"
tail -3 gpurun_out/bench_c2_${N}gpu.err
nvidia-smi topo -m 2>/dev/null | head -14
