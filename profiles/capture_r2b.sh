#!/bin/bash
# Round 2, last session: the rigidbody2d pipeline and the ball2d portal branch (no ncu evidence until now), 500 000 bodies each, on ONE B200:
#   r2_time_<tag>.json  CUDA-event timing per kernel (not under a profiler); r2_launches_<tag>.csv  launch list; r2_full_<tag>.csv  ncu --set full raw page
# profiles/summarize_r2.py c6 c7 turns them into profiles/ncu_r2_{c6,c7}.{json,md}.
set -u
mkdir -p gpurun_out
for cfg in 6 7; do
  tag=c$cfg
  python profiles/ncu_target.py --config $cfg --n 500000 --time --steps 10 --warmup 3 > gpurun_out/r2_time_$tag.json 2> gpurun_out/r2_time_$tag.err
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$tag.csv python profiles/ncu_target.py --config $cfg --n 500000 --steps 1 --warmup 1 > /dev/null 2>&1
  timeout 70 ncu --set full --clock-control none -k regex:"sg_|k_ball2d|k_rb2d|k_b2p|k_r2p" -c 60 -f -o gpurun_out/r2_full_$tag python profiles/ncu_target.py --config $cfg --n 500000 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_full_$tag.err
  ncu -i gpurun_out/r2_full_$tag.ncu-rep --page raw --csv > gpurun_out/r2_full_$tag.csv 2>/dev/null
  rm -f gpurun_out/r2_full_$tag.ncu-rep
  tail -c 400 gpurun_out/r2_time_$tag.json
done
ls -la gpurun_out | grep "r2_.*c[67]"
