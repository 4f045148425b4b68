mkdir -p gpurun_out/r02q
timeout 600 python -m pytest tests/test_gpu_xgemm.py tests/test_gpu_train.py -x -q -m gpu -s -k "tc2 or layouts or bf16_engine" > gpurun_out/r02q/pytest.txt 2>&1; echo rc=$? >> gpurun_out/r02q/pytest.txt
grep -n "tc2 engine\|passed\|failed\|Error\|assert" gpurun_out/r02q/pytest.txt | head -20
for e in tc tc2; do for w in cfg4 cfg5; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --f32-engine $e > gpurun_out/r02q/${w}_$e.json 2> gpurun_out/r02q/${w}_$e.err; tail -c 300 gpurun_out/r02q/${w}_$e.err
done; done
python - <<'P'
import json
for e in ('tc','tc2'):
  for f in ('cfg4','cfg5'):
    for line in open('gpurun_out/r02q/%s_%s.json'%(f,e)):
        if line.startswith('{'):
            d=json.loads(line); print(e, f, d['value'], d['ms_per_step'], d.get('loss_first_step'), d.get('loss_last_step'), d.get('gpu_launches'))
P
