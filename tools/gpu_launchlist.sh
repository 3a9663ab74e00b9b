#!/bin/bash
# per-kernel launch lists (ncu, durations only) of the HS+tracer step at T170 L40 and of the moist model at T85 L40
mkdir -p gpurun_out
SMALL="python bench.py --steps 4 --warmup 3 --spinup 20 --no-cpu-baseline --e2e-steps 1 --no-moist"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_moist.csv python tools/moistbench.py T85 40 360 5 --ncu-steps 3 > gpurun_out/ncu_moist.log 2>&1
ls -la gpurun_out
