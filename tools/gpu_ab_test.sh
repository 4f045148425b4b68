# parity tests of the inference path + cfg3 bench with alternative builds (scratch_ab/<variant>.so); restores the product build
# usage: gpu_ab_test.sh <outdir> <variant>...
out=gpurun_out/$1; shift; mkdir -p $out
cp transferable3d_b200/libt3d_b200.so /tmp/base.so
for v in "$@" base; do
  if [ $v = base ]; then cp /tmp/base.so transferable3d_b200/libt3d_b200.so; else cp scratch_ab/$v.so transferable3d_b200/libt3d_b200.so; fi
  if [ $v != base ]; then
    timeout 240 python -m pytest tests/test_gpu_bench_path.py tests/test_gpu_parity.py -x -q -m gpu > $out/pytest_$v.txt 2>&1; rc=$?
    echo "$v pytest rc=$rc: $(tail -1 $out/pytest_$v.txt)"
    if [ $rc -ne 0 ]; then continue; fi
  fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-exact-mode > $out/cfg3_$v.json 2> $out/cfg3_$v.err || tail -c 400 $out/cfg3_$v.err
  python - $out/cfg3_$v.json $v <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        print(sys.argv[2], 'value %.0f ms %.3f e2e %.0f | stage2 %.3f ms frac %.3f | seg1 %.3f ms frac %.3f | clocks %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'],
              d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_fused_maxpool']['avg_launch_ms'], d['roofline_fused_maxpool']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']),
              'parity', {k: (v.get('mask_point_agreement'), v.get('seg_logit_err_max_of_scale')) for k, v in d.get('parity', {}).items() if isinstance(v, dict)})
P
done
cp /tmp/base.so transferable3d_b200/libt3d_b200.so
