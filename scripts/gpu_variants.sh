(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) | tee gpurun_out/pytest_gpu.log
python bench_ops.py > gpurun_out/ops.json 2> gpurun_out/ops.err
python -c "
import json
d=json.load(open('gpurun_out/ops.json'))
for k,v in d['ops'].items():
    if 'affine' in k or 'interp' in k: print('%-48s %.4f ms  %.3f' % (k, v['ms'], v['frac_of_hbm_peak']))
"
