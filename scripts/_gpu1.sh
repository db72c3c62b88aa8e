timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_dist.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline']['step_share'])"
