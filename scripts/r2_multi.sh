#!/bin/bash
# round-2 multi-GPU run on a box with G GPUs: copy-bandwidth sweep and bench.py at N = 1, 2, 4, ... G
G=${1:-2}
mkdir -p gpurun_out
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) "$@"; }
: > gpurun_out/r2_h2d_sweep.jsonl
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  TR $n scripts/h2d_sweep.py 2> gpurun_out/r2_h2d_${n}.err | grep '^{' >> gpurun_out/r2_h2d_sweep.jsonl
done
cat gpurun_out/r2_h2d_sweep.jsonl
for n in 1 2 4 8; do
  [ $n -le $G ] || continue
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/r2_scale_${n}gpu.json 2> gpurun_out/r2_scale_${n}gpu.err
  else TR $n bench.py --gpus $n > gpurun_out/r2_scale_${n}gpu.json 2> gpurun_out/r2_scale_${n}gpu.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_scale_${n}gpu.json").read().strip().splitlines()[-1])
    a = d.get("also", {})
    print("N=%d value %.2f G frac %.3f e2e %.2f G | c3_256 %.2f G | c3_atlas %s subjects/s %.1f ms (%s) numa %s %s" % (
        d["n_gpus"], d["value"] / 1e9, d["hbm_roofline_frac_96B"], d["e2e"]["value"] / 1e9, a.get("c3_256", {}).get("value", 0) / 1e9,
        a.get("c3_atlas", {}).get("value"), a.get("c3_atlas", {}).get("ms_per_epoch", 0), a.get("c3_atlas", {}).get("error"),
        d["config"].get("host_numa_node"), d["config"].get("host_numa_note")))
except Exception as e:
    print("N=${n} ERR", e); print(open("gpurun_out/r2_scale_${n}gpu.err").read()[-1500:])
PY
done
grep -h "NCCL INFO.*\(nranks\|NVLS\|Connected all\|comm 0x\)" gpurun_out/r2_scale_*gpu.err | head -20
