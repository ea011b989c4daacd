(timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python __graft_entry__.py --smoke 2>&1 | tail -1
