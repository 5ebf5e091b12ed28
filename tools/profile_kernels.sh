#!/bin/bash
# On the GPU box: key ncu metrics of every kernel of the secondary configurations (one pass each).
# usage: tools/profile_kernels.sh <tag> [configs_once.py config, default all] [log2 particles]
D=gpurun_out/prof2
mkdir -p $D
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,lts__t_bytes.sum,smsp__warps_eligible.avg.per_cycle_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
ncu --metrics $M --clock-control none --csv --log-file $D/$1_config_kernels.csv python tools/configs_once.py ${2:-all} ${3:-24} > $D/$1_configs_once.log 2>&1
ls -la $D/$1_*
