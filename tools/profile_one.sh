#!/bin/bash
# On the GPU box: one full ncu capture of one kernel of one configuration, kept as CSV pages.
# usage: tools/profile_one.sh <tag> <kernel regex> <configs_once.py config> [log2 particles]
D=gpurun_out/prof2
mkdir -p $D
ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -o $D/$1 python tools/configs_once.py $3 ${4:-24} > /dev/null 2>&1
ncu -i $D/$1.ncu-rep --page raw --csv > $D/$1_raw.csv 2>/dev/null
ncu -i $D/$1.ncu-rep --page source --csv > $D/$1_source.csv 2>/dev/null
rm -f $D/$1.ncu-rep
ls -la $D/$1_*
