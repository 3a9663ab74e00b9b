#!/bin/bash
# First GPU visit of round 2: everything that was built after the last GPU visit of round 1 (profiles/README.md, "What has no GPU
# evidence yet"), in the order cheapest / most informative first.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
# 1. the new GPU tests, one file at a time, no -x inside a file: one failing option must not hide the others
( time timeout 1500 python -m pytest tests/test_gpu_rows_f.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02a_pytest_rows_f.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_rrtm.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02a_pytest_rrtm.txt 2>&1
# 2. memcheck of the kernels that have never run (small sizes; the tests above already compare their results)
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rows_f.py -m gpu -q -x \
  -k "hs_forcing_options_parity or betts_miller_parity or seasonal_insolation or do_read_co2 or dry_convection_parity" \
  > gpurun_out/r02a_sanitizer_rows_f.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/rrtm_bench.py 32 16 40 1 > gpurun_out/r02a_sanitizer_rrtm.txt 2>&1
# 3. the established suite and the bench (regression check of round 1's numbers on the new library build)
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_physics.py tests/test_gpu_moist.py -m gpu -q -x ) > gpurun_out/r02a_pytest_established.txt 2>&1
timeout 600 python bench.py --steps 300 --warmup 10 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
# 4. RRTMG: first timings, launch list and full captures of the two kernels (tools/gpu_rrtm.sh without its pytest / sanitizer part)
timeout 300 python tools/physbench.py > gpurun_out/r02a_physbench.json 2> gpurun_out/r02a_physbench.err
timeout 300 python tools/rrtm_bench.py > gpurun_out/r02a_rrtm_bench.json 2> gpurun_out/r02a_rrtm_bench.err
timeout 420 python tools/mima_bench.py > gpurun_out/r02a_mima_bench.json 2> gpurun_out/r02a_mima_bench.err
SMALL="python tools/rrtm_bench.py 256 128 40 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02a_rrtm_launches.csv $SMALL > gpurun_out/ncu_rrtm_launch.log 2>&1
for k in rrtmg_lw_kernel rrtmg_sw_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02a_prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
