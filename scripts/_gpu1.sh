timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python bench.py --steps 10 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -3 gpurun_out/bench_r1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['config'], d['e2e']['value'], d['gpu_launches'], d['roofline']['step_share'])"
