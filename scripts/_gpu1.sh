timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu -k "graph" 2>&1 | tail -8
python scripts/run_strips.py --native --reps 200 --check --graph 2>&1 | grep -E '^\{|rror' | cut -c1-330
