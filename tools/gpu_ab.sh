# A/B of alternative builds of libt3d_b200.so (scratch_ab/*.so) on the cfg3 bench: usage gpu_ab.sh <outdir> <variant>...
out=gpurun_out/$1; shift
mkdir -p $out
cp transferable3d_b200/libt3d_b200.so /tmp/base.so
for v in base "$@"; do
  if [ $v = base ]; then cp /tmp/base.so transferable3d_b200/libt3d_b200.so; else cp scratch_ab/$v.so transferable3d_b200/libt3d_b200.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 $BENCH_FLAGS > $out/cfg3_$v.json 2> $out/cfg3_$v.err || tail -c 400 $out/cfg3_$v.err
  python - $out/cfg3_$v.json $v <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        print(sys.argv[2], 'value %.0f ms %.3f e2e %.0f | stage2 %.3f ms frac %.3f | seg1 %.3f ms frac %.3f | clocks %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'],
              d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_fused_maxpool']['avg_launch_ms'], d['roofline_fused_maxpool']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']),
              'parity', {k: v.get('mask_point_agreement') for k, v in d.get('parity', {}).items() if isinstance(v, dict)})
P
done
cp /tmp/base.so transferable3d_b200/libt3d_b200.so
