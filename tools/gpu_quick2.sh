#!/bin/bash
# quick visit: GPU tests + short bench (no CPU arm, no moist arm) + launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-moist --e2e-steps 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
SMALL="python bench.py --steps 4 --warmup 3 --spinup 20 --no-cpu-baseline --e2e-steps 1 --no-moist"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/pytest_gpu.txt
