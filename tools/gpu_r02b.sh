#!/bin/bash
# Round 2, visit B: GPU suite after the visit-A fixes, RRTMG timings + full captures of the reworked kernels.
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -rf --timeout 600 -s -k "not multirank" ) > gpurun_out/r02b_pytest_gpu.txt 2>&1
timeout 300 python tools/rrtm_bench.py > gpurun_out/r02b_rrtm_bench.json 2> gpurun_out/r02b_rrtm_bench.err
SMALL="python tools/rrtm_bench.py 256 128 40 2"
for k in rrtmg_lw_kernel rrtmg_sw_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02b_prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
for f in gpurun_out/*.ncu-rep; do
  if [ $(stat -c %s "$f") -gt 25000000 ]; then ncu -i "$f" --page raw --csv > "${f%.ncu-rep}_raw.csv" 2>/dev/null; rm -f "$f"; fi
done
du -sh gpurun_out
grep -v "^  (" gpurun_out/r02b_pytest_gpu.txt | tail -15
cat gpurun_out/r02b_rrtm_bench.json
