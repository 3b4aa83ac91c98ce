#!/bin/bash
# HISTORICAL: SG_BP_NAR / SG_BP_PASS1 were development knobs (operand planes, kernel variants); the current tree ignores them.
set -u
python -m pytest tests/test_ball2d_gpu.py tests/test_slab_gpu.py tests/test_multi_gpu.py tests/test_rb2d_gpu.py tests/test_portals_gpu.py tests/test_config1_gpu.py -m gpu -x -q 2>&1 | tail -2
for nar in 0 1; do
for v in 1 2; do
  for c in "--config 2 --steps 20" "--config 3 --steps 10" "--config 3 --n 16777216 --steps 5"; do
    echo "nar $nar variant $v $c: $(SG_BP_NAR=$nar SG_BP_PASS1=$v python profiles/ncu_target.py $c --time --warmup 3 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["us"] for k,v in d["kernels"].items() if k in ("bp_count","bp_side","bp_emit","bp_contacts","bp_scatter")})')"
  done
done
done
