mkdir -p gpurun_out/r02p
timeout 900 python -m pytest tests/test_gpu_lazy_bn.py tests/test_gpu_train.py tests/test_gpu_semisup_train.py tests/test_gpu_semisup_a_train.py tests/test_gpu_boxpc_variants.py tests/test_gpu_xgemm.py -x -q -m gpu > gpurun_out/r02p/pytest.txt 2>&1; echo rc=$? >> gpurun_out/r02p/pytest.txt
tail -5 gpurun_out/r02p/pytest.txt
python tools/gpu_gemm_bn_bench.py | tee gpurun_out/r02p/gemm_bn.txt
for w in cfg4 cfg5; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02p/$w.json 2> gpurun_out/r02p/$w.err; tail -c 300 gpurun_out/r02p/$w.err
done
python - <<'P'
import json
for f in ('cfg4','cfg5'):
    for line in open('gpurun_out/r02p/%s.json'%f):
        if line.startswith('{'):
            d=json.loads(line); print(f, d['value'], d['ms_per_step'], d.get('loss_first_step'), d.get('loss_last_step'), d.get('gpu_launches'))
P
