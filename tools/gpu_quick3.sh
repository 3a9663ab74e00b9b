#!/bin/bash
# GPU tests + bench (with moist arm, no CPU arm) + moist launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
python bench.py --steps 300 --warmup 10 --no-cpu-baseline --e2e-steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_moist.csv python tools/moistbench.py T85 40 360 5 --ncu-steps 3 > gpurun_out/ncu_moist.log 2>&1
tail -5 gpurun_out/pytest_gpu.txt
