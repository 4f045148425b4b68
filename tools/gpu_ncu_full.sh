# ncu --set full of the dominant kernels of the final build (1 GPU) + launch lists.  usage: bash tools/gpu_ncu_full.sh
set -x
mkdir -p gpurun_out/r02n
timeout 600 python -m pytest tests/test_gpu_lazy_bn.py tests/test_gpu_semisup_train.py -x -q -m gpu > gpurun_out/r02n/pytest.txt 2>&1; echo rc=$? >> gpurun_out/r02n/pytest.txt
tail -5 gpurun_out/r02n/pytest.txt
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02n/cfg5.json 2> gpurun_out/r02n/cfg5.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02n/cfg5.json | head -1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,sm__cycles_elapsed.avg.per_second,sm__inst_executed.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
timeout 900 ncu --metrics $M --clock-control none -k 'regex:seg_stage2|chain_max' -c 14 --csv --log-file gpurun_out/r02n/ncu_cfg3_kernels.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r02n/cfg3_full.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k 'regex:xgemm|xg_pp|xg_as|pool_bn|maxpool|colstats|bn_backward' -s 40 -c 40 --csv --log-file gpurun_out/r02n/ncu_cfg4_kernels.csv \
    python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02n/cfg4_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02n/launches_cfg5.csv python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02n/ncu_cfg5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n/launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02n/ncu_cfg4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02n/launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/r02n/ncu_cfg3.log 2>&1
rm -f gpurun_out/r02n/*.log
du -sh gpurun_out/r02n
