#!/bin/bash
# Round 2, visit R (1 GPU): the default bench line of the final state (convection / condensation post kernels timed as their own groups)
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err
tail -3 gpurun_out/r02r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench.json'))
print("ms_per_step", d["ms_per_step"], "steady", d["steady_state"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "rrtmg_call_ms", d["config"]["rrtmg_call_ms"], "clocks", d["clocks"], "launches", d["gpu_launches"])
print({k: round(v,4) for k,v in d["kernel_groups_ms"].items()})
print("roofline", d["roofline"])
for k,v in d["kernels"].items(): print(k, v)
print("cpu", d["cpu_baseline"])
for k,v in d["extra"].items(): print(k, v.get("ms_per_step"), v.get("error"))
PY
