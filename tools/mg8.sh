set -x
mkdir -p gpurun_out/r02c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/r02c/topo.txt 2>&1
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02c/cfg3_weak_n8.json 2> gpurun_out/r02c/cfg3_weak_n8.err
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-parity --scaling strong > gpurun_out/r02c/cfg3_strong_n8.json 2> gpurun_out/r02c/cfg3_strong_n8.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg4 > gpurun_out/r02c/cfg4_n8.json 2> gpurun_out/r02c/cfg4_n8.err
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/r02c/cfg5_n8.json 2> gpurun_out/r02c/cfg5_n8.err
tail -c 600 gpurun_out/r02c/*.json
