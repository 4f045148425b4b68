# ncu metric tables + launch list of the cfg3 kernels of the final build (1 GPU), micro-benchmarks.  usage: bash tools/gpu_ncu_final.sh <outdir>
out=gpurun_out/$1; mkdir -p $out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,sm__cycles_elapsed.avg.per_second,sm__inst_executed.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
timeout 900 ncu --metrics $M --clock-control none -k 'regex:seg_stage2|chain_max' -c 14 --csv --log-file $out/ncu_cfg3_kernels.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/cfg3_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $out/ncu_cfg3.log 2>&1
timeout 60 ./scratch_ab/ldtm_bw > $out/ubench_ldtm_bw.txt 2>&1
timeout 60 ./scratch_ab/ldtm_mma > $out/ubench_ldtm_mma.txt 2>&1
timeout 120 python tools/gpu_trace.py > $out/trace.txt 2>&1
ENGINES=tc2 NEV=40 timeout 120 python tools/gpu_trace_xgemm.py > $out/trace_xgemm_tc2.txt 2>&1
rm -f $out/*.log; ls -la $out
