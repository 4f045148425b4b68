# full GPU validation: pytest -m gpu, smoke(), default bench line; outputs under gpurun_out/$1
out=gpurun_out/$1; mkdir -p $out
timeout 1500 python -m pytest tests/ -x -q -m gpu > $out/pytest.txt 2>&1; echo "pytest rc=$?" >> $out/pytest.txt; tail -3 $out/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?" >> $out/smoke.txt; tail -3 $out/smoke.txt
timeout 600 python bench.py > $out/bench_cfg3.json 2> $out/bench_cfg3.err; echo "bench rc=$?"; tail -c 300 $out/bench_cfg3.err
python - $out/bench_cfg3.json <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line)
        print('value %.0f ms %.3f e2e %.0f | stage2 %.3f ms frac %.3f | seg1 %.3f ms frac %.3f | clocks %s %s' % (d['value'], d['ms_per_step'], d['e2e']['value'],
              d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_fused_maxpool']['avg_launch_ms'], d['roofline_fused_maxpool']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
        print('exact', d.get('exact_mode', {}).get('value'), 'cpu', d.get('cpu_baseline'))
P
