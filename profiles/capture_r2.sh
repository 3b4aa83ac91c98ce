#!/bin/bash
# Round 2 captures on ONE B200 (gpurun --timeout 1500 -- 'bash profiles/capture_r2.sh'):
#   per config (2: 1M pile, 3: 2M and 16M gas, 4: 4.1M rb3d spheres split_ham, 5: mixed sphere/box/mesh)
#     r2_time_<tag>.json      resident-step timing with CUDA events per kernel (NOT under a profiler)
#     r2_launches_<tag>.csv   ncu --metrics gpu__time_duration.sum launch list of one warm step
#     r2_full_<tag>.csv       ncu --set full raw page of the same step's kernels (not for the 16M scene: ncu's save/restore of 10 GB per replay)
# profiles/summarize_r2.py turns them into profiles/ncu_r2_<tag>.{json,md}.
set -u
mkdir -p gpurun_out
run() {
  tag=$1; shift
  python profiles/ncu_target.py "$@" --time --steps 10 --warmup 3 > gpurun_out/r2_time_$tag.json 2> gpurun_out/r2_time_$tag.err
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$tag.csv python profiles/ncu_target.py "$@" --steps 1 --warmup 1 > /dev/null 2>&1
}
full() {
  tag=$1; shift
  ncu --set full --clock-control none -k regex:"sg_|k_ball2d|k_rb3d|k_slab" -c 90 -f -o gpurun_out/r2_full_$tag python profiles/ncu_target.py "$@" --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_full_$tag.err
  ncu -i gpurun_out/r2_full_$tag.ncu-rep --page raw --csv > gpurun_out/r2_full_$tag.csv 2>/dev/null
  rm -f gpurun_out/r2_full_$tag.ncu-rep
}
run c2 --config 2;            full c2 --config 2
run c3_2m --config 3;         full c3_2m --config 3
run c3_16m --config 3 --n 16777216
run c4 --config 4;            full c4 --config 4
run c5 --config 5;            full c5 --config 5
ls -la gpurun_out | grep r2_ 
