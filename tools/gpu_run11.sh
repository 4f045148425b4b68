mkdir -p gpurun_out/r02z
timeout 300 python -m pytest tests/test_gpu_xgemm.py -x -q -m gpu -k "tc2" > gpurun_out/r02z/pytest_xgemm.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02z/pytest_xgemm.txt
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_lazy_bn.py -x -q -m gpu -s -k "tc2 or lazy" > gpurun_out/r02z/pytest_train.txt 2>&1; echo "rc=$?"; grep -n "tc2 engine\|passed\|failed" gpurun_out/r02z/pytest_train.txt | cut -c 1-600
for w in cfg4 cfg5; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --f32-engine tc2 > gpurun_out/r02z/${w}_tc2.json 2> gpurun_out/r02z/${w}_tc2.err; tail -c 300 gpurun_out/r02z/${w}_tc2.err
done
python - <<'P'
import json
for f in ('cfg4','cfg5'):
    for line in open('gpurun_out/r02z/%s_tc2.json'%f):
        if line.startswith('{'):
            d=json.loads(line); print(f, d['value'], d['ms_per_step'], d.get('loss_first_step'), d.get('loss_last_step'), d.get('gpu_launches'))
P
