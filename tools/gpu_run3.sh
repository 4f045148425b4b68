set -x
mkdir -p gpurun_out/r02g
timeout 900 python -m pytest tests/test_gpu_boxpc_variants.py -q -m gpu > gpurun_out/r02g/pytest_variants.txt 2>&1; echo rc=$? >> gpurun_out/r02g/pytest_variants.txt
tail -60 gpurun_out/r02g/pytest_variants.txt
