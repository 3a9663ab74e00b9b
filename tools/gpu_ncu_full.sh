#!/bin/bash
# full ncu captures of selected kernels (NCU_KERNELS, regex per kernel) from a short bench run
mkdir -p gpurun_out
SMALL="python bench.py --steps 4 --warmup 3 --spinup 20 --no-cpu-baseline --e2e-steps 1 --no-moist"
for k in ${NCU_KERNELS:-tracer_horiz_kernel tracer_ppm_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
