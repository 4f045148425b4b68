set -x
mkdir -p gpurun_out/r02f
timeout 900 python -m pytest tests/test_gpu_lazy_bn.py tests/test_gpu_train.py tests/test_gpu_semisup_train.py -x -q -m gpu > gpurun_out/r02f/pytest.txt 2>&1; echo rc=$? >> gpurun_out/r02f/pytest.txt
tail -12 gpurun_out/r02f/pytest.txt
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/cfg4.json 2> gpurun_out/r02f/cfg4.err; tail -c 400 gpurun_out/r02f/cfg4.err
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/cfg5.json 2> gpurun_out/r02f/cfg5.err; tail -c 400 gpurun_out/r02f/cfg5.err
python - <<'P'
import json
for f in ('cfg4','cfg5'):
    for line in open('gpurun_out/r02f/%s.json'%f):
        if line.startswith('{'):
            d=json.loads(line); print(f, d['value'], d['ms_per_step'], d.get('loss_first_step'), d.get('loss_last_step'), d.get('gpu_launches'))
P
