#!/bin/bash
# Round 2, visit F (1 GPU): RRTMG occupancy variants, GPU suite with the staged Legendre epilogue / reordered forward grid, headline bench
mkdir -p gpurun_out
bash tools/rrtm_variants.sh run > gpurun_out/r02f_rrtm_variants.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -k "not multirank" ) > gpurun_out/r02f_pytest_gpu.txt 2>&1
timeout 1200 python bench.py --steps 96 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
du -sh gpurun_out
cat gpurun_out/r02f_rrtm_variants.txt
tail -6 gpurun_out/r02f_pytest_gpu.txt
tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print("ms_per_step", d["ms_per_step"], "steady", d["steady_state"]["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print({k: round(v,4) for k,v in d["kernel_groups_ms"].items()})
for k,v in d["extra"].items(): print(k, v.get("ms_per_step"), v.get("error"))
PY
