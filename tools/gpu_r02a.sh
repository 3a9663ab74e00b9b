#!/bin/bash
# Round 2, visit A: the whole GPU suite without -x (every failure visible), memcheck of the never-run kernels, first RRTMG / physics
# timings and full ncu captures of the RRTMG and the column-physics kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( which gfortran mpif90 flang nvfortran; nproc; lscpu | grep -i "model name"; nvidia-smi -L ) > gpurun_out/r02a_box.txt 2>&1
( time timeout 1700 python -m pytest tests -m gpu -q -rf --timeout 600 ) > gpurun_out/r02a_pytest_gpu.txt 2>&1
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rows_f.py -m gpu -q -x \
  -k "hs_forcing_options_parity or betts_miller_parity or seasonal_insolation or do_read_co2 or dry_convection_parity" \
  > gpurun_out/r02a_sanitizer_rows_f.txt 2>&1
timeout 400 compute-sanitizer --tool memcheck python tools/rrtm_bench.py 32 16 40 1 > gpurun_out/r02a_sanitizer_rrtm.txt 2>&1
timeout 300 python tools/physbench.py > gpurun_out/r02a_physbench.json 2> gpurun_out/r02a_physbench.err
timeout 300 python tools/rrtm_bench.py > gpurun_out/r02a_rrtm_bench.json 2> gpurun_out/r02a_rrtm_bench.err
timeout 420 python tools/mima_bench.py > gpurun_out/r02a_mima_bench.json 2> gpurun_out/r02a_mima_bench.err
SMALL="python tools/rrtm_bench.py 256 128 40 2"
for k in rrtmg_lw_kernel rrtmg_sw_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02a_prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
# column physics at T170 through the moist model: launch list of 3 steps after a 1-day spin-up, then one full capture of one whole step
MOIST="python tools/moistbench.py T170 40 150 1 --ncu-steps 3"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_moist_launches.csv $MOIST > gpurun_out/ncu_moist_launch.log 2>&1
MOIST1="python tools/moistbench.py T170 40 150 1 --ncu-steps 1"
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -f -o gpurun_out/r02a_prof_moist_step $MOIST1 > gpurun_out/ncu_moist_full.log 2>&1
# gpurun_out/ must stay below 64 MiB or nothing is copied back: turn any large report into its raw csv page and drop it
for f in gpurun_out/*.ncu-rep; do
  if [ $(stat -c %s "$f") -gt 25000000 ]; then ncu -i "$f" --page raw --csv > "${f%.ncu-rep}_raw.csv" 2>/dev/null; rm -f "$f"; fi
done
du -sh gpurun_out
tail -25 gpurun_out/r02a_pytest_gpu.txt
cat gpurun_out/r02a_box.txt gpurun_out/r02a_rrtm_bench.json gpurun_out/r02a_mima_bench.json
