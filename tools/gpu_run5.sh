mkdir -p gpurun_out/r02i
timeout 600 python tools/gpu_mask_exactness.py > gpurun_out/r02i/mask_exactness.txt 2>&1
timeout 600 python tools/gpu_mode_throughput.py > gpurun_out/r02i/mode_throughput.txt 2>&1
timeout 900 python bench.py > gpurun_out/r02i/bench_cfg3.json 2> gpurun_out/r02i/bench_cfg3.err
tail -5 gpurun_out/r02i/mask_exactness.txt; tail -8 gpurun_out/r02i/mode_throughput.txt; tail -c 300 gpurun_out/r02i/bench_cfg3.err
python - <<'P'
import json
for line in open('gpurun_out/r02i/bench_cfg3.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['e2e']['value'], json.dumps(d.get('parity'))[:1500]); print(json.dumps(d.get('modes'))[:800])
P
