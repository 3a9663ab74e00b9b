#!/bin/bash
# Round 2, visit C: the four tests that failed in visit B, RRTMG timings, first run of the MiMA T170 L40 headline bench with the extra configurations
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_baseline_parity.py tests/test_gpu_physics.py tests/test_gpu_rrtm.py tests/test_gpu_rows_f.py -m gpu -q -rf --timeout 600 -s ) > gpurun_out/r02c_pytest_gpu.txt 2>&1
timeout 300 python tools/rrtm_bench.py > gpurun_out/r02c_rrtm_bench.json 2> gpurun_out/r02c_rrtm_bench.err
timeout 900 python bench.py --steps 96 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
du -sh gpurun_out
grep -v "^  (" gpurun_out/r02c_pytest_gpu.txt | tail -12
grep -A40 "T170 L40: (step" gpurun_out/r02c_pytest_gpu.txt | head -60
cat gpurun_out/r02c_rrtm_bench.json
tail -5 gpurun_out/r02c_bench.err
cat gpurun_out/r02c_bench.json
