#!/bin/bash
# First GPU call of round 2 (ONE B200):  gpurun --timeout 900 -- 'bash profiles/first_call_r2.sh'
# Runs what round 1 could only check on the CPU, file by file so that one failure does not hide the rest, then the full suite,
# the bench line and the portal timings.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
for f in tests/test_zw_rb3d_minertia_gpu.py tests/test_zx_rb2d_resident_gpu.py tests/test_zy_portal_trajectory_gpu.py tests/test_zz_rb2d_portals_gpu.py tests/test_zz_rb3d_portals_gpu.py; do
  b=$(basename $f .py)
  timeout 240 python -m pytest $f -q -m gpu > gpurun_out/$b.log 2>&1
  echo "$b: $(tail -1 gpurun_out/$b.log)"
done
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "full suite: $(tail -1 gpurun_out/pytest_gpu_full.log)"
timeout 240 python bench.py > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err; tail -c 600 gpurun_out/bench_r2_first.json
for sysm in ball2d rb2d rb3d; do
  timeout 90 python profiles/portal_timing.py $sysm 1048576 > gpurun_out/portal_timing_${sysm}_r2.json 2> gpurun_out/portal_timing_${sysm}.err
  echo "$sysm: $(head -c 400 gpurun_out/portal_timing_${sysm}_r2.json) $(tail -2 gpurun_out/portal_timing_${sysm}.err)"
done
