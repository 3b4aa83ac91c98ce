#!/bin/bash
# HISTORICAL: the SG_BP_PASS1 knob (tma | 0..3) existed only in the development commits of sg_bp_count_l1 (git log: "Pass 1 (2-D) without staging");
# the numbers it produced are quoted in DESIGN.md 4.2.1 / 4.4.  In the current tree the variable is ignored.
# A/B of pass 1 (2-D): SG_BP_PASS1 = tma (staged, warp-specialised) | 0..3 (unstaged variants: chunk / min blocks per SM); parity first
set -u
mkdir -p gpurun_out
python -m pytest tests/test_ball2d_gpu.py tests/test_slab_gpu.py tests/test_multi_gpu.py tests/test_rb2d_gpu.py tests/test_portals_gpu.py tests/test_config1_gpu.py -m gpu -x -q 2>&1 | tail -4
for v in tma 0 1 2 3; do
  for c in "--config 2 --steps 20" "--config 3 --steps 10" ; do
    echo "variant $v $c: $(SG_BP_PASS1=$v python profiles/ncu_target.py $c --time --warmup 3 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["us"] for k,v in d["kernels"].items() if k in ("bp_count","bp_emit","bp_contacts","bp_scatter")})')"
  done
done
for v in tma 0 1; do
  echo "variant $v 16M: $(SG_BP_PASS1=$v python profiles/ncu_target.py --config 3 --n 16777216 --steps 5 --time --warmup 2 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["us"] for k,v in d["kernels"].items()})')"
done
