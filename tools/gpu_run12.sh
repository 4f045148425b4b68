mkdir -p gpurun_out/r03g
timeout 400 python -m pytest tests/test_gpu_bench_path.py tests/test_gpu_parity.py tests/test_gpu_dataset.py tests/test_gpu_postprocess.py -x -q -m gpu > gpurun_out/r03g/pytest.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r03g/pytest.txt
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-exact-mode > gpurun_out/r03g/cfg3.json 2> gpurun_out/r03g/cfg3.err || tail -c 800 gpurun_out/r03g/cfg3.err
python - <<'P'
import json
for line in open('gpurun_out/r03g/cfg3.json'):
    if line.startswith('{'):
        d=json.loads(line); e=d['e2e']
        print('value %.0f ms %.3f | e2e %.0f ms %.3f h2d %d d2h %d launches %s' % (d['value'], d['ms_per_step'], e['value'], e['ms_per_step'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e.get('gpu_launches')))
        print(json.dumps(d.get('parity', {}).get('bf16', {}))[:600])
P
