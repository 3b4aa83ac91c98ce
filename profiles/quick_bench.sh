timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/b.json; python - <<PY
import json
d=json.load(open("gpurun_out/b.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"])
print({k:round(v["ms_per_step"]*1e3,1) for k,v in d["roofline"]["kernels"].items()})
PY
