#!/bin/bash
# Round 2, visit H (1 GPU): column-per-lane longwave kernel: timings of the register-limit variants vs the g-point kernel, GPU suite, bench
mkdir -p gpurun_out
bash tools/rrtm_variants.sh run > gpurun_out/r02h_rrtm_variants.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -k "not multirank" ) > gpurun_out/r02h_pytest_gpu.txt 2>&1
timeout 1200 python bench.py --steps 96 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
SMALL="python tools/rrtm_bench.py 512 256 40 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rrtmg_lw_col_kernel -s 1 -c 1 -f -o gpurun_out/r02h_prof_rrtmg_lw_col_kernel $SMALL > gpurun_out/ncu_lwc.log 2>&1
du -sh gpurun_out
cat gpurun_out/r02h_rrtm_variants.txt
tail -6 gpurun_out/r02h_pytest_gpu.txt
tail -3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print("ms_per_step", d["ms_per_step"], "steady", d["steady_state"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "rrtmg_call_ms", d["config"]["rrtmg_call_ms"])
print({k: round(v,4) for k,v in d["kernel_groups_ms"].items()})
PY
