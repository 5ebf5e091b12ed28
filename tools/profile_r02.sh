#!/bin/bash
# On the GPU box: the ncu evidence of round 2.  Writes gpurun_out/prof2/ (scratch, CSV pages only: the .ncu-rep files are
# ~26 MB each and gpurun_out/ travels back only below 64 MiB); tools/summarise_r02.py turns it into profiles/.
# usage: tools/profile_r02.sh <tag>
set -x
D=gpurun_out/prof2
mkdir -p $D
T=${1:-r02}
# 1. launch list of the bench command itself (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/${T}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --cpu-particles 1000 --c5-file-particles 100000 --file-particles 1000000 > $D/${T}_bench_under_ncu.log 2>&1
# 2. every kernel of every secondary configuration once, key metrics
bash tools/profile_kernels.sh $T all 24 > /dev/null
# 3. full captures with source-level counters: headline kernel (2e8 particles), staged C3 / C4 / C5, row kernel of C5 (emitting runs)
ncu --set full --clock-control none --import-source on -k regex:k_sis_fused -s 1 -c 1 -o $D/${T}_fused_c2 \
    python bench.py --steps 1 --warmup 3 --particles 200000000 --cpu-particles 1000 --no-configs --strong-particles 1000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sis_staged -c 1 -o $D/${T}_staged_c3 python tools/configs_once.py c3 24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sis_staged -c 1 -o $D/${T}_staged_c4 python tools/configs_once.py c4 24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sis_staged -c 1 -o $D/${T}_staged_c5 python tools/configs_once.py c5 24 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sis_rows -c 1 -o $D/${T}_rows_c5 python tools/configs_once.py c5rows 24 > /dev/null 2>&1
for n in fused_c2 staged_c3 staged_c4 staged_c5 rows_c5; do
    ncu -i $D/${T}_$n.ncu-rep --page raw --csv > $D/${T}_${n}_raw.csv 2>/dev/null
    ncu -i $D/${T}_$n.ncu-rep --page source --csv > $D/${T}_${n}_source.csv 2>/dev/null
    rm -f $D/${T}_$n.ncu-rep
done
ls -la $D
