#!/bin/bash
# Round 2, visit I (1 GPU): column-per-lane LW (rebalanced bands, g-points in halves) and SW kernels: register-limit variants vs the
# g-point kernels, RRTMG GPU tests on the default library, one ncu capture of each column kernel on a small batch
mkdir -p gpurun_out
bash tools/rrtm_variants.sh run > gpurun_out/r02i_rrtm_variants.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_rrtm.py tests/test_gpu_rows_f.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02i_pytest_rrtm.txt 2>&1
SMALL="python tools/rrtm_bench.py 512 256 40 1"
for k in rrtmg_lw_col_kernel rrtmg_sw_col_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02i_prof_$k $SMALL > gpurun_out/ncu_$k.log 2>&1
done
du -sh gpurun_out
cat gpurun_out/r02i_rrtm_variants.txt
tail -6 gpurun_out/r02i_pytest_rrtm.txt
