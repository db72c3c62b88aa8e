CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:ofb:: --csv --log-file gpurun_out/r1_launches_final.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:lk_level_kernelILi9ELi2ELb0E -s 3 -c 1 -f -o gpurun_out/r1_lk_level0_final $CMD > gpurun_out/ncu2.log 2>&1
echo "set full rc=$?"
tail -3 gpurun_out/ncu2.log
