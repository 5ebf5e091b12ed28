#!/bin/bash
# On the GPU box: ncu evidence for round 2.  Writes gpurun_out/prof2/ (scratch); tools/summarise_r02.py turns it into profiles/.
set -x
D=gpurun_out/prof2
mkdir -p $D
T=${1:-r02}
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,lts__t_bytes.sum,smsp__warps_eligible.avg.per_cycle_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum
# every kernel of every secondary configuration once, key metrics
ncu --metrics $M --clock-control none --csv --log-file $D/${T}_config_kernels.csv python tools/configs_once.py all 24 > $D/${T}_configs_once.log 2>&1
# full captures of the staged kernel (C3 and C4) with source-level counters
ncu --set full --clock-control none --import-source on -k regex:k_sis_staged -c 1 -o $D/${T}_staged_c3 python tools/configs_once.py c3 24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sis_staged -c 1 -o $D/${T}_staged_c4 python tools/configs_once.py c4 24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sis_rows -c 1 -o $D/${T}_rows_c5 python tools/configs_once.py c5 24 > /dev/null 2>&1
# the reports are ~26 MB each and gpurun_out/ travels back only below 64 MiB: keep their CSV pages instead
for n in staged_c3 staged_c4 rows_c5; do
    ncu -i $D/${T}_$n.ncu-rep --page raw --csv > $D/${T}_${n}_raw.csv 2>/dev/null
    ncu -i $D/${T}_$n.ncu-rep --page source --csv > $D/${T}_${n}_source.csv 2>/dev/null
    rm -f $D/${T}_$n.ncu-rep
done
ls -la $D
