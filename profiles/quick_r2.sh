#!/bin/bash
# parity of the broad-phase users, then resident-step timings of configs 2, 3 (2M and 16M) and 4
set -u
python -m pytest tests/test_ball2d_gpu.py tests/test_slab_gpu.py tests/test_multi_gpu.py tests/test_rb2d_gpu.py tests/test_rb3d_gpu.py tests/test_rb3d_slab_gpu.py tests/test_portals_gpu.py tests/test_config1_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
for c in "--config 2 --steps 20" "--config 3 --steps 10" "--config 3 --n 16777216 --steps 5" "--config 4 --steps 5"; do
  echo "$c: $(python profiles/ncu_target.py $c --time --warmup 3 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["us"] for k,v in d["kernels"].items()})')"
done
