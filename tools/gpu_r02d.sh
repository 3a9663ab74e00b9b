#!/bin/bash
# Round 2, visit D (1 GPU): whole GPU suite after the one-all-reduce restructure of the global sums, RRTMG timings, headline bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -rf --timeout 600 -s -k "not multirank" ) > gpurun_out/r02d_pytest_gpu.txt 2>&1
timeout 300 python tools/rrtm_bench.py > gpurun_out/r02d_rrtm_bench.json 2> gpurun_out/r02d_rrtm_bench.err
timeout 1200 python bench.py --steps 96 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
du -sh gpurun_out
grep -v "^  (" gpurun_out/r02d_pytest_gpu.txt | tail -8
grep -B2 -A26 "T170 L40: (step" gpurun_out/r02d_pytest_gpu.txt | head -40
cat gpurun_out/r02d_rrtm_bench.json
tail -5 gpurun_out/r02d_bench.err
cat gpurun_out/r02d_bench.json
