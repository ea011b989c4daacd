python -m pytest tests/test_shoot_gpu.py -x -q -m gpu -k "ring" 2>&1 | tail -3
for r in 1 2 3; do
LGM_ADSTAR_RING_256=0 LGM_RING_256=2 python scripts/variant_bench.py c3
LGM_RING_256=2 python scripts/variant_bench.py c3
python scripts/variant_bench.py c3
done
