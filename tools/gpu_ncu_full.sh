# ncu --set full of the dominant kernels of the final build (1 GPU).  usage: bash tools/gpu_ncu_full.sh
set -x
mkdir -p gpurun_out/r02n
# cfg3 bf16 + f16x2 legs: one resident step each after one warm-up (ncu replays each captured launch ~40 times)
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:seg_stage2|chain_max' -c 14 -o gpurun_out/r02n/cfg3_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r02n/cfg3_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:xgemm|xg_pp|xg_as|pool_bn|maxpool|colstats|bn_backward' -s 60 -c 40 -o gpurun_out/r02n/cfg4_full \
    python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02n/cfg4_full.log 2>&1
ls -la gpurun_out/r02n
