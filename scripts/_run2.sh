for r in 1 2 3 4; do
LGM_ADSTAR_RING_256=0 python scripts/variant_bench.py c3
python scripts/variant_bench.py c3
done
