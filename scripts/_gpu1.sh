timeout 300 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 900 gpurun_out/bench_r1.json; tail -3 gpurun_out/bench_r1.err
timeout 300 python bench.py --pairs 1024 --steps 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
timeout 300 python bench.py --pairs 512 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
