# cfg4 / cfg5 bench (tc2) with alternative builds: usage gpu_ab_train.sh <outdir> <variant>...
out=gpurun_out/$1; shift; mkdir -p $out
cp transferable3d_b200/libt3d_b200.so /tmp/base.so
for v in "$@" base; do
  if [ $v = base ]; then cp /tmp/base.so transferable3d_b200/libt3d_b200.so; else cp scratch_ab/$v.so transferable3d_b200/libt3d_b200.so; fi
  for w in cfg4 cfg5; do
    timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --f32-engine tc2 > $out/${w}_$v.json 2> $out/${w}_$v.err
    python - $out/${w}_$v.json $v $w <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[2], sys.argv[3], round(d['value']), round(d['ms_per_step'],3), d['loss_last_step'])
P
  done
done
cp /tmp/base.so transferable3d_b200/libt3d_b200.so
