#!/bin/bash
# On the GPU box: ncu evidence for the round.  Writes gpurun_out/prof/*.csv (+ .ncu-rep for the headline kernel).
set -x
mkdir -p gpurun_out/prof
R=${1:-r01}
# 1. launch list of the default bench command (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/prof/${R}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --cpu-particles 1000 > gpurun_out/prof/${R}_bench_under_ncu.log 2>&1
# 2. full capture of the dominant kernel (2e8 particles, one launch)
ncu --set full --clock-control none --import-source on -k regex:k_sis_fused -s 1 -c 1 -o gpurun_out/prof/${R}_k_sis_fused \
    python bench.py --steps 1 --warmup 3 --particles 200000000 --cpu-particles 1000 > /dev/null 2>&1
# 3. row path: C5 (hmm 1000 steps) and C3 (linear gaussian 32): every kernel once, key metrics
cat > /tmp/rows_once.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import analytic
from cpprob_b200 import Engine
g = analytic.golden()
with Engine(seed=0x5EED) as e:
    e.run("hmm", g["obs_hmm_1000"], 1 << 20)
    e.run("hmm", g["obs_hmm_64"], 1 << 24)
    e.run("linear_gaussian_1d", g["obs_linear_gaussian_32"], 1 << 24)
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_bytes.sum
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/prof/${R}_rows_kernels.csv python /tmp/rows_once.py > /dev/null 2>&1
ls -la gpurun_out/prof
