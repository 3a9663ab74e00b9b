#!/bin/bash
# Round 2, visit U (1 GPU): the default bench line and the reference arm with the all-C++ CPU baseline, as the driver runs them
mkdir -p gpurun_out
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u_bench_reference.json 2> gpurun_out/r02u_bench_reference.err
timeout 1500 python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
tail -3 gpurun_out/r02u_bench_reference.err; cut -c1-200 gpurun_out/r02u_bench_reference.json
python - <<'PY'
import json
r=json.load(open('gpurun_out/r02u_bench_reference.json'))
print("reference ms_per_step", r["ms_per_step"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["parts_s"])
d=json.load(open('gpurun_out/r02u_bench.json'))
print("ms_per_step", d["ms_per_step"], "steady", d["steady_state"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "clocks", d["clocks"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["ms_per_step"], d["cpu_baseline"]["parts_s"])
for k,v in d["extra"].items(): print(k, v.get("ms_per_step"), v.get("error"))
PY
tail -3 gpurun_out/r02u_bench.err
