#!/bin/bash
# Run under gpurun on ONE B200:  gpurun --timeout 900 -- 'bash profiles/capture.sh r1'
# 1. the bench line (not under a profiler), 2. the ncu launch list of the same command, 3. one `--set full` capture of the
# hot kernels.  Outputs land in gpurun_out/; profiles/summarize.py turns them into the tracked summaries.
set -u
R=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${R}_1gpu.json 2> gpurun_out/bench_${R}_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_${R}_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sg_bp_(emit|count|contacts|scatter|hist)|k_ball2d_prep" -s 12 -c 6 -o gpurun_out/prof_${R} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -8
