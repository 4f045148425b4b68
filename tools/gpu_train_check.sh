# training-path check: GEMM / lazy-BN / training-step tests, then cfg4 / cfg5 bench for both engines.  usage: gpu_train_check.sh <outdir>
out=gpurun_out/$1; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_xgemm.py tests/test_gpu_lazy_bn.py tests/test_gpu_train.py tests/test_gpu_semisup_train.py tests/test_gpu_semisup_a_train.py tests/test_gpu_boxpc_variants.py -x -q -m gpu > $out/pytest_train.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_train.txt
for e in tc2 tc; do for w in cfg4 cfg5; do
  timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --f32-engine $e > $out/${w}_$e.json 2> $out/${w}_$e.err
  python - $out/${w}_$e.json $e $w <<'P'
import json,sys
for line in open(sys.argv[1]):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[2], sys.argv[3], round(d['value']), round(d['ms_per_step'],3), d['loss_first_step'], d['loss_last_step'])
P
done; done
