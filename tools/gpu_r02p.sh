#!/bin/bash
# Round 2, final visit P (1 GPU): smoke, the whole GPU suite, the default bench line (extras + CPU baseline) and the reference arm,
# launch list of the bench command's timed region, full captures of its top kernels in their final form
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02p_smoke.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 ) > gpurun_out/r02p_pytest_gpu.txt 2>&1
timeout 1500 python bench.py > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02p_bench_reference.json 2> gpurun_out/r02p_bench_reference.err
SMALL="python bench.py --steps 48 --warmup 3 --spinup 96 --no-extra --no-cpu-baseline --e2e-steps 1 --profile-range"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02p_launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
for k in rrtmg_lw_col_kernel sbm_convection_kernel vert_diff_down_kernel; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r02p_prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
for f in gpurun_out/*.ncu-rep; do
  if [ $(stat -c %s "$f") -gt 12000000 ]; then ncu -i "$f" --page raw --csv > "${f%.ncu-rep}_raw.csv" 2>/dev/null; rm -f "$f"; fi
done
du -sh gpurun_out
cat gpurun_out/r02p_smoke.txt | tail -5
tail -6 gpurun_out/r02p_pytest_gpu.txt
tail -3 gpurun_out/r02p_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_bench.json'))
print("ms_per_step", d["ms_per_step"], "steady", d["steady_state"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "rrtmg_call_ms", d["config"]["rrtmg_call_ms"], "clocks", d["clocks"])
print({k: round(v,4) for k,v in d["kernel_groups_ms"].items()})
print("roofline", d["roofline"])
print("cpu", d["cpu_baseline"])
for k,v in d["extra"].items(): print(k, v.get("ms_per_step"), v.get("error"), v.get("cpu_baseline"))
PY
cut -c1-900 gpurun_out/r02p_bench_reference.json; tail -2 gpurun_out/r02p_bench_reference.err
