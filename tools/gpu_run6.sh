mkdir -p gpurun_out/r02j
timeout 900 python -m pytest tests/test_gpu_lazy_bn.py tests/test_gpu_semisup_train.py -x -q -m gpu > gpurun_out/r02j/pytest.txt 2>&1; echo rc=$? >> gpurun_out/r02j/pytest.txt
tail -6 gpurun_out/r02j/pytest.txt
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02j/cfg5.json 2> gpurun_out/r02j/cfg5.err; tail -c 300 gpurun_out/r02j/cfg5.err
python - <<'P'
import json
for line in open('gpurun_out/r02j/cfg5.json'):
    if line.startswith('{'):
        d=json.loads(line); print('cfg5', d['value'], d['ms_per_step'], d.get('loss_first_step'), d.get('loss_last_step'), d.get('gpu_launches'))
P
