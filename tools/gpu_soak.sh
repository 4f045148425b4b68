# soak: repeated bench runs under per-run timeouts; a hang shows as rc=124.  usage: gpu_soak.sh <outdir>
out=gpurun_out/$1; mkdir -p $out
for i in $(seq 1 14); do
  timeout 90 python bench.py --workload cfg5 --steps 20 --warmup 3 --no-cpu-baseline --f32-engine tc2 > $out/cfg5_tc2_$i.json 2> $out/cfg5_tc2_$i.err; echo "cfg5 tc2 run $i rc=$?"
done
for i in $(seq 1 8); do
  timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > $out/cfg3_$i.json 2> $out/cfg3_$i.err; echo "cfg3 run $i rc=$?"
done
grep -h -o '"ms_per_step": [0-9.]*' $out/cfg3_*.json | sort | uniq -c | head -30
