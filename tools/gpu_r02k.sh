#!/bin/bash
# Round 2, visit K (1 GPU): LW column kernel with 128-bit evict-last table reads and 2 / 4 terms in flight: register-limit variants;
# RRTMG + rows-f GPU tests on the default library (setcoef pre-pass scratch fix)
mkdir -p gpurun_out
bash tools/rrtm_variants.sh run > gpurun_out/r02k_rrtm_variants.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_rrtm.py tests/test_gpu_rows_f.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02k_pytest_rrtm.txt 2>&1
SMALL="python tools/rrtm_bench.py 512 256 40 1"
ISCA_B200_LIB=$PWD/isca_b200/lib/variants/u2_5/libisca_b200.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:rrtmg_lw_col_kernel -s 1 -c 1 -f -o gpurun_out/r02k_prof_rrtmg_lw_col_kernel $SMALL > gpurun_out/ncu_lwc.log 2>&1
du -sh gpurun_out
cat gpurun_out/r02k_rrtm_variants.txt
tail -6 gpurun_out/r02k_pytest_rrtm.txt
