mkdir -p gpurun_out/r02h
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02h/pytest_all.txt 2>&1; echo rc=$? >> gpurun_out/r02h/pytest_all.txt
tail -15 gpurun_out/r02h/pytest_all.txt
