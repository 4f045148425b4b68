mkdir -p gpurun_out/r02r
for e in tc2; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r/launches_cfg4_$e.csv python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu-baseline --f32-engine $e > gpurun_out/r02r/ncu_$e.log 2>&1
done
tail -2 gpurun_out/r02r/ncu_tc2.log | cut -c 1-300
