timeout 600 python -m pytest tests/test_affine_atlas_gpu.py -x -q 2>&1 | tail -8
python bench_affine_atlas.py --steps 2 --warmup 1 2>&1 | tail -2
