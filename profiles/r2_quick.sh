#!/bin/bash
# quick loop: 2-D parity tests, then resident-step timings of configs 2 and 3 (2M and 16M)
set -u
mkdir -p gpurun_out
python -m pytest tests/test_ball2d_gpu.py tests/test_slab_gpu.py tests/test_multi_gpu.py tests/test_rb2d_gpu.py tests/test_portals_gpu.py -m gpu -x -q 2>&1 | tail -8
python profiles/ncu_target.py --config 2 --time --steps 20 --warmup 3 2>&1 | tail -1
python profiles/ncu_target.py --config 3 --time --steps 10 --warmup 3 2>&1 | tail -1
python profiles/ncu_target.py --config 3 --n 16777216 --time --steps 5 --warmup 2 2>&1 | tail -1
