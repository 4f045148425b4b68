mkdir -p gpurun_out/r03b
for i in 1 2 3 4 5 6; do
  for w in cfg5 cfg4; do
    timeout 90 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --f32-engine tc2 > gpurun_out/r03b/${w}_$i.json 2> gpurun_out/r03b/${w}_$i.err; echo "$w run $i rc=$?"
  done
done
nvidia-smi --query-gpu=temperature.gpu,clocks.sm --format=csv
