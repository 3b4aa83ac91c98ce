#!/bin/bash
# round 2, first GPU call: where does the time go on configs 3 and 4 (ncu --set full of one whole step each), config 3 at full size on one GPU
set -u
mkdir -p gpurun_out
python profiles/ncu_target.py --config 3 --time --steps 10 --warmup 3 > gpurun_out/t_c3_2m.json 2> gpurun_out/t_c3_2m.err
python profiles/ncu_target.py --config 3 --n 16777216 --time --steps 5 --warmup 2 > gpurun_out/t_c3_16m.json 2> gpurun_out/t_c3_16m.err
python profiles/ncu_target.py --config 4 --time --steps 5 --warmup 2 > gpurun_out/t_c4.json 2> gpurun_out/t_c4.err
python profiles/ncu_target.py --config 5 --time --steps 5 --warmup 2 > gpurun_out/t_c5.json 2> gpurun_out/t_c5.err
ncu --set full --clock-control none --import-source on -k regex:"sg_|k_ball2d|k_rb3d" -c 40 -o gpurun_out/prof_r2a_c3 python profiles/ncu_target.py --config 3 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/ncu_c3.err
ncu --set full --clock-control none --import-source on -k regex:"sg_|k_ball2d|k_rb3d" -c 60 -o gpurun_out/prof_r2a_c4 python profiles/ncu_target.py --config 4 --steps 1 --warmup 1 > /dev/null 2> gpurun_out/ncu_c4.err
cat gpurun_out/t_c3_2m.json gpurun_out/t_c3_16m.json gpurun_out/t_c4.json gpurun_out/t_c5.json
ls -la gpurun_out | tail -12
