#!/bin/bash
# One GPU-box visit: GPU tests, smoke, both bench arms, ncu launch list and full captures of the top kernels.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 500 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
SMALL="python bench.py --steps 4 --warmup 3 --spinup 20 --no-cpu-baseline --e2e-steps 1 --no-moist"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launch.log 2>&1
rm -f gpurun_out/*.ncu-rep
for k in ${NCU_KERNELS:-fft_inv_kernel grid_step_sigma_kernel tracer_horiz_kernel tracer_ppm_kernel legendre_inv_kernel legendre_fwd_kernel fft_fwd_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
