#!/bin/bash
# First GPU visit of the RRTMG row (a30): its GPU tests, the two informational bench tools, the launch list of one radiation call
# and full ncu captures of the LW and SW kernels on a T85-sized batch.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_rrtm.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_rrtm.py -m gpu -q ) 2>&1 | tail -40 > gpurun_out/pytest_gpu_rrtm.txt
timeout 300 python tools/rrtm_bench.py > gpurun_out/rrtm_bench.json 2> gpurun_out/rrtm_bench.err
timeout 300 python tools/rrtm_bench.py 256 128 40 > gpurun_out/rrtm_bench_t85.json 2>> gpurun_out/rrtm_bench.err
timeout 420 python tools/mima_bench.py > gpurun_out/mima_bench.json 2> gpurun_out/mima_bench.err
SMALL="python tools/rrtm_bench.py 256 128 40 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/rrtm_launches.csv $SMALL > gpurun_out/ncu_rrtm_launch.log 2>&1
for k in rrtmg_lw_kernel rrtmg_sw_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
timeout 600 compute-sanitizer --tool memcheck python tools/rrtm_bench.py 32 16 40 1 > gpurun_out/sanitizer_rrtm.txt 2>&1
ls -la gpurun_out
