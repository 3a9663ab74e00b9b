#!/bin/bash
# Round 2, visit J (1 GPU): LW column kernel with the setcoef pre-pass (one evaluation per (column, layer) instead of one per band)
# and the term-list builder as one non-inlined function: register-limit variants; RRTMG GPU tests; ncu capture of the pre_4 variant
mkdir -p gpurun_out
bash tools/rrtm_variants.sh run > gpurun_out/r02j_rrtm_variants.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_rrtm.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02j_pytest_rrtm.txt 2>&1
SMALL="python tools/rrtm_bench.py 512 256 40 1"
ISCA_B200_LIB=$PWD/isca_b200/lib/variants/pre_4/libisca_b200.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:rrtmg_lw_col_kernel -s 1 -c 1 -f -o gpurun_out/r02j_prof_rrtmg_lw_col_kernel $SMALL > gpurun_out/ncu_lwc.log 2>&1
ISCA_B200_LIB=$PWD/isca_b200/lib/variants/pre_4/libisca_b200.so timeout 600 ncu --clock-control none -k regex:rrtmg_lw_setcoef_kernel -s 1 -c 1 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum $SMALL > gpurun_out/r02j_ncu_setcoef.txt 2>&1
du -sh gpurun_out
cat gpurun_out/r02j_rrtm_variants.txt
tail -6 gpurun_out/r02j_pytest_rrtm.txt
grep -E "gpu__time|dram__bytes" gpurun_out/r02j_ncu_setcoef.txt
