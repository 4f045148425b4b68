# ncu launch list of one training workload: usage gpu_launches.sh <outdir> <workload cfg4|cfg5> <engine tc|tc2>
mkdir -p gpurun_out/$1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$1/launches_$2_$3.csv python bench.py --workload $2 --steps 2 --warmup 3 --no-cpu-baseline --f32-engine $3 > gpurun_out/$1/ncu_$2_$3.log 2>&1
tail -1 gpurun_out/$1/ncu_$2_$3.log | cut -c 1-200
