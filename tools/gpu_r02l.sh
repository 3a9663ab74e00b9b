#!/bin/bash
# Round 2, visit L (1 GPU): column-physics kernels with 7 resident CTAs per SM (72 registers) against the previous register counts,
# vert_diff_down / surface_flux variants; RRTMG timing of the default library (LW column kernel + setcoef pre-pass, 5 CTAs/SM);
# physics + moist + RRTMG GPU tests
mkdir -p gpurun_out
bash tools/phys_variants.sh run > gpurun_out/r02l_phys_variants.txt 2>&1
python tools/rrtm_bench.py > gpurun_out/r02l_rrtm_bench.json 2> gpurun_out/r02l_rrtm_bench.err
( time timeout 1200 python -m pytest tests/test_gpu_physics.py tests/test_gpu_moist.py tests/test_gpu_rrtm.py tests/test_gpu_rows_f.py -m gpu -q -rf --timeout 600 ) > gpurun_out/r02l_pytest.txt 2>&1
du -sh gpurun_out
cat gpurun_out/r02l_phys_variants.txt
python -c "
import json; d=json.load(open('gpurun_out/r02l_rrtm_bench.json')); print('lw_ms', d['lw_kernel_ms'], 'sw_ms', d['sw_kernel_ms'])"
tail -6 gpurun_out/r02l_pytest.txt
