set -x
mkdir -p gpurun_out/r02e
timeout 600 python -m pytest tests/test_gpu_lazy_bn.py -x -q -m gpu > gpurun_out/r02e/pytest_lazy.txt 2>&1; echo rc=$? >> gpurun_out/r02e/pytest_lazy.txt
tail -30 gpurun_out/r02e/pytest_lazy.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e/launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02e/ncu_cfg4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02e/launches_cfg5.csv python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02e/ncu_cfg5.log 2>&1
