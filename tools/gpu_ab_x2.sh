# A/B of an alternative build on the f16x2 (exact) mode: x2 parity tests + cfg3 bench in f16x2 precision.  usage: gpu_ab_x2.sh <outdir> <variant>
out=gpurun_out/$1; v=$2; mkdir -p $out
cp transferable3d_b200/libt3d_b200.so /tmp/base.so
cp scratch_ab/$v.so transferable3d_b200/libt3d_b200.so
timeout 300 python -m pytest tests/test_gpu_x2.py tests/test_gpu_bench_path.py -x -q -m gpu -k "x2 or f16x2" > $out/pytest_$v.txt 2>&1; rc=$?
echo "$v pytest rc=$rc: $(tail -1 $out/pytest_$v.txt)"
if [ $rc -eq 0 ]; then
for w in $v base; do
  if [ $w = base ]; then cp /tmp/base.so transferable3d_b200/libt3d_b200.so; fi
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --precision f16x2 --no-exact-mode > $out/cfg3_x2_$w.json 2> $out/cfg3_x2_$w.err || tail -c 400 $out/cfg3_x2_$w.err
  python - $out/cfg3_x2_$w.json $w <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        print(sys.argv[2], 'value %.0f ms %.3f | stage2 %.3f ms | seg1 %.3f ms | clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline_fused_maxpool']['avg_launch_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons']),
              'parity', {k: (v.get('mask_point_agreement'), v.get('mask_frustums_bit_exact')) for k, v in d.get('parity', {}).items() if isinstance(v, dict)})
P
done
fi
cp /tmp/base.so transferable3d_b200/libt3d_b200.so
