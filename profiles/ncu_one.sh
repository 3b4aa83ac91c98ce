#!/bin/bash
# ncu --set full of ONE kernel of one config, exported on the box to CSV (raw metrics + per-source-line), the report kept if small:
#   bash profiles/ncu_one.sh <tag> <kernel regex> <skip> <ncu_target args...>
set -u
TAG=$1; KRE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG python profiles/ncu_target.py "$@" > /dev/null 2> gpurun_out/ncu_$TAG.err
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page details > gpurun_out/ncu_${TAG}_details.txt 2>/dev/null
ls -la gpurun_out/prof_$TAG.ncu-rep
if [ $(stat -c %s gpurun_out/prof_$TAG.ncu-rep) -gt 30000000 ]; then rm gpurun_out/prof_$TAG.ncu-rep; fi
