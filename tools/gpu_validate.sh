#!/bin/bash
# One short GPU-box visit: GPU tests, smoke, bench, launch list (no full captures).
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --steps 500 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/devbench.py > gpurun_out/devbench_default.json 2> gpurun_out/devbench_default.err
SMALL="python bench.py --steps 4 --warmup 3 --spinup 4 --no-cpu-baseline --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out
